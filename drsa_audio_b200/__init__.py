"""drsa_audio_b200 -- B200 (sm_100a) kernels for the DRSA / LRP explanation hot path.

This package only holds what the path needs: the CUDA sources (``csrc/``), the built
C-ABI library (``libdrsa_b200.so``, built in-tree by ``csrc/build.sh``) and the ctypes
binding (``_lib``).  The user-facing API is the ``cxai`` package, which mirrors the
reference's ``cxai.xai`` names.
"""
from ._lib import lib, check, build, library_path, DRSAError, PREC_FP32, PREC_TC_F16X2  # noqa: F401

__all__ = ["lib", "check", "build", "library_path", "DRSAError", "PREC_FP32", "PREC_TC_F16X2"]
