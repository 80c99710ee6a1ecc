"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol that
include/drsa_b200.h declares, and refuses to compute without an sm_100 device (no fallback)."""
import ctypes
import os
import re

import pytest

import drsa_audio_b200
from drsa_audio_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "drsa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:drsa|lrp|logmel)_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(drsa_audio_b200.library_path()):
        drsa_audio_b200.build()
    return drsa_audio_b200.lib()


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/drsa_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in drsa_audio_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_status_strings(lib):
    assert lib.drsa_version() == 100
    assert lib.drsa_status_string(0) == b"ok"
    assert b"sm_100" in lib.drsa_status_string(-3)


def test_workspace_queries_need_no_gpu(lib):
    assert lib.drsa_step_workspace_bytes(16000, 64, 64, 4, _lib.PREC_FP32) > 0
    assert lib.drsa_step_workspace_bytes(640000, 256, 256, 4, _lib.PREC_TC_F16X2) > 0
    assert lib.drsa_step_workspace_bytes(640000, 64, 64, 4, _lib.PREC_TC_F16X2) == -2     # shape unsupported
    assert lib.drsa_step_workspace_bytes(0, 64, 64, 4, _lib.PREC_FP32) == -1
    assert lib.drsa_step_workspace_bytes(100, 64, 63, 4, _lib.PREC_FP32) == -1            # m % K != 0
    assert lib.drsa_finish_workspace_bytes(256, 256) > 0


def test_peer_exchange_and_fused_pool_queries_need_no_gpu(lib):
    """Shape / argument checks of the entry points added for the multi-GPU exchange and the fused conv + pool."""
    n = lib.drsa_exchange_bytes(256, 256, 4, 8)
    assert n == 64 * 4 + 2 * 8 * (256 * 256 + 64) * 8                   # header + 2 parities x 8 ranks x padded d*m+K {value, flag} words
    assert lib.drsa_exchange_bytes(256, 256, 4, 1) == -1                # a single rank has nothing to exchange
    assert lib.drsa_exchange_bytes(256, 256, 4, 9) == -2                # more than DRSA_MAX_PEERS
    assert lib.drsa_exchange_bytes(48, 48, 4, 2) == -2                  # not a shape of the fused finish kernel
    px = _lib.PeerExchange()
    px.world, px.rank = 1, 0
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    st = lib.drsa_finish_step_p2p(ctypes.byref(px), p, 100, p, 64, 64, 4, None, None, None, None, 0, 8, 1e-6, 0, None, p, 64, None)
    assert st == -1
    # 8 x 16 tiles or whole 8 x 8 maps, power-of-two windows that fit a warp's rows
    assert lib.lrp_tc_conv3x3_pool_supported(64, 64, 64, 128, 256, 2, 4) == 0
    assert lib.lrp_tc_conv3x3_pool_supported(3, 128, 128, 8, 8, 2, 2) == 0
    assert lib.lrp_tc_conv3x3_pool_supported(3, 64, 64, 4, 4, 2, 2) == -2
    assert lib.lrp_tc_conv3x3_pool_supported(3, 64, 64, 32, 32, 3, 3) == -2
    assert lib.lrp_tc_conv3x3_pool_supported(3, 64, 64, 32, 32, 4, 2) == -2         # 4 rows do not fit a warp of a 16-wide tile


def test_compute_fails_loudly_without_sm100(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    st = lib.drsa_step(p, p, p, None, None, 8, 4, 4, 2, 0, 1.0, 1.0, 1.0, p, p, 1 << 20, None)
    assert st == -3
    with pytest.raises(_lib.DRSAError):
        _lib.check(st, "drsa_step")


def test_python_mirror_refuses_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    with pytest.raises(_lib.DRSAError):
        SubspaceOptimizer(torch.eye(8), torch.rand(16, 8), torch.rand(16, 8), None, num_concepts=2, device="cpu")
    with pytest.raises(_lib.DRSAError):
        SubspaceOptimizer(torch.eye(8), torch.rand(16, 8), torch.rand(16, 8), None, num_concepts=2, device="cuda")


def test_pad_plan_for_shapes_off_the_tensor_core_grid(lib):
    """Zero-padding plans of cxai.xai.drsa.drsa._pad_plan (host logic, no GPU): (d', m', d_k')."""
    from cxai.xai.drsa.drsa import _pad_plan
    assert _pad_plan(100, 100, 4, 200000) == (128, 128, 32)        # arch A layer 19: 4 x 25 -> 4 x 32
    assert _pad_plan(64, 64, 4, 200000) == (128, 128, 32)
    assert _pad_plan(96, 96, 2, 20000) == (128, 128, 64)
    assert _pad_plan(200, 200, 4, 100000) == (256, 256, 64)
    assert _pad_plan(100, 100, 5, 100000) is None                  # 5 x 32 = 160 columns: not a multiple of 128
    assert _pad_plan(100, 100, 3, 100000) is None                  # m % K != 0


def test_pow2_scale():
    from cxai.xai.drsa.drsa import _pow2_scale
    for mx in (1e-6, 0.03, 0.25, 1.0, 77.0, 5e4, 3e9):
        s = _pow2_scale(mx)
        assert 128.0 <= mx * s < 256.0
        assert s == 2.0 ** round(__import__("math").log2(s))
    assert _pow2_scale(0.0) == 1.0


def test_integration_notes_cover_every_declared_entry_point():
    """INTEGRATION.md names, for every entry point of include/drsa_b200.h, the reference code it replaces (or says that there
    is none)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "drsa_b200.h")).read()
    notes = open(os.path.join(root, "INTEGRATION.md")).read()
    declared = sorted(set(re.findall(r"\b((?:drsa|lrp|logmel)_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 60
    # `drsa_ipc_alloc/open/close/free` is written as one entry in the table
    missing = [s for s in declared if s not in notes and not (s.startswith("drsa_ipc_") and "drsa_ipc_alloc/open/close/free" in notes)]
    assert not missing, missing
