"""ctypes binding of libdrsa_b200.so (declarations follow include/drsa_b200.h 1:1).

There is deliberately no fallback: if the library has not been built, or the device is
not an sm_100 part, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
PREC_FP32 = 0
PREC_TC_F16X2 = 1
PREC_TC_F16 = 2
PREC_TC_F16_AC2 = 3
PREC_TC_F32C = 4


class DRSAError(RuntimeError):
    pass


def library_path() -> str:
    return os.path.join(HERE, "libdrsa_b200.so")


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    script = os.path.join(HERE, "csrc", "build.sh")
    res = subprocess.run(["bash", script], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise DRSAError("building libdrsa_b200.so failed")
    return library_path()


_i64, _i32, _f32, _vp = C.c_int64, C.c_int, C.c_float, C.c_void_p

# name -> (restype, argtypes); mirrors include/drsa_b200.h
SIGNATURES = {
    "drsa_status_string": (C.c_char_p, [_i32]),
    "drsa_version": (_i32, []),
    "drsa_last_cuda_error": (_i32, []),
    "drsa_check_device": (_i32, [_i32]),
    "drsa_pack_f16": (_i32, [_vp, _i64, _f32, _vp, _vp]),
    "drsa_pack_f16_hilo": (_i32, [_vp, _i64, _f32, _vp, _vp, _vp]),
    "drsa_absmax": (_i32, [_vp, _i64, _vp, _vp]),
    "drsa_step_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _i32]),
    "drsa_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _vp, _i64, _vp]),
    "drsa_rownorm_max": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "drsa_split_u": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "drsa_sums_combine": (_i32, [_vp, _vp, _f32, _vp, _i64, _vp]),
    "drsa_qr_retract": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "drsa_finish_step_qr": (_i32, [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "drsa_finish_workspace_bytes": (_i64, [_i32, _i32]),
    "drsa_finish_step": (_i32, [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _vp, _vp,
                                _i64, _vp]),
    "drsa_polar_retract": (_i32, [_vp, _i32, _i32, _vp, _i32, _f32, _vp, _vp, _i64, _vp]),
    "drsa_subspace_relevances": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "drsa_subspace_relevances_workspace_bytes": (_i64, [_i64, _i64, _i32, _i32]),
    "drsa_subset_objectives": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "drsa_context_pairs_nhwc": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "drsa_context_gather": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "drsa_context_vectors": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "drsa_sumsq": (_i32, [_vp, _i64, _vp, _vp]),
    "drsa_normalize": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp]),
    "lrp_conv3x3_forward": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lrp_conv3x3_flip_weights": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "lrp_conv3x3_backward": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp]),
    "lrp_maxpool_forward": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "lrp_maxpool_backward": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lrp_dense_forward": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "lrp_dense_epsilon_backward": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp]),
    "lrp_relu_mask": (_i32, [_vp, _vp, _i64, _vp]),
    "lrp_tc_conv3x3_supported": (_i32, [_i64, _i32, _i32, _i32, _i32]),
    "lrp_tc_conv3x3_forward": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp,
                                      _vp, _vp]),
    "lrp_tc_conv3x3_first": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "lrp_tc_maxpool": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "lrp_tc_maxpool_backward": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lrp_tc_conv3x3_ratio": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp,
                                    _vp]),
    "lrp_tc_conv3x3_inputmul": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                       _vp]),
    "lrp_tc_sample_absmax_ratio": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "lrp_tc_relu_mask": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "lrp_tc_planes_to_f32": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "lrp_subspace_filter_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "lrp_subspace_project": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "lrp_subspace_filter": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _i64, _vp]),
    "lrp_tc_nhwc_f32_to_nchw": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lrp_tc_nchw_to_nhwc_f32": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lrp_tc_nhwc_to_nchw": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lrp_tc_split_f16": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "logmel_transform_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "logmel_transform_wav": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _i64,
                                    _vp]),
    "drsa_selftest_umma": (_i32, [_i32, _vp]),
    "drsa_debug_set_tc_profile": (_i32, [_vp]),
    "drsa_debug_tc_kernel_attrs": (_i32, [_i32, _i32, _vp]),
    "drsa_debug_set_tc_variant": (_i32, [_i32]),
    "lrp_debug_set_conv_variant": (_i32, [_i32]),
}

MAX_PEERS = 8


class PeerExchange(C.Structure):
    """drsa_peer_exchange of include/drsa_b200.h."""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("buffers", C.c_void_p * MAX_PEERS)]


SIGNATURES.update({
    "drsa_exchange_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "drsa_finish_step_p2p": (_i32, [C.POINTER(PeerExchange), _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64,
                                    _i32, _f32, _i32, _vp, _vp, _i64, _vp]),
    "lrp_tc_first_ones_backward": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "lrp_tc_conv3x3_pool_supported": (_i32, [_i64, _i32, _i32, _i32, _i32, _i32, _i32]),
    "lrp_tc_conv3x3_forward_pool": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                           _vp, _vp, _vp, _vp, _vp]),
    "drsa_ipc_alloc": (_i32, [_i64, C.POINTER(C.c_void_p), C.c_char_p]),
    "drsa_ipc_open": (_i32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "drsa_ipc_close": (_i32, [_vp]),
    "drsa_ipc_free": (_i32, [_vp]),
})

_lock = threading.Lock()
_lib = None


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = library_path()
                if not os.path.isfile(path):
                    raise DRSAError(
                        f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU or PyTorch fallback for this path)")
                handle = C.CDLL(path)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def status_string(status: int) -> str:
    return lib().drsa_status_string(int(status)).decode()


def check(status: int, what: str = "") -> int:
    """Raise DRSAError for a negative status; returns non-negative values unchanged."""
    if status < 0:
        l = lib()
        msg = l.drsa_status_string(int(status)).decode()
        extra = f" (cudaError {l.drsa_last_cuda_error()})" if status == -6 else ""
        raise DRSAError(f"{what or 'libdrsa_b200'}: {msg}{extra}")
    return status
