"""What the shipped library is built on, read from its SASS (no GPU needed; cuobjdump comes with the toolkit):
the row-pass and convolution kernels issue tcgen05 MMAs fed by TMA with accumulators in tensor memory, nothing goes through
the legacy mma.sync / wgmma paths, and the peer exchange of the finish kernels polls 8-byte words with system-scope loads and
pushes 16-byte words (value and flag of a word must not be torn apart -- csrc/retract_fused.cu: peer_push / peer_poll)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
SO = os.path.join(ROOT, "drsa_audio_b200", "libdrsa_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("cu++filt") is None or not os.path.isfile(SO),
                                reason="needs cuobjdump, cu++filt and the built library")


@pytest.fixture(scope="module")
def sass():
    import sass_summary
    return sass_summary.summarise(SO)


def test_hot_kernels_use_tcgen05_tma_and_tensor_memory(sass):
    row_pass = [k for k in sass if k.startswith("drsa_tc_step_kernel<")]
    convs = [k for k in sass if k.startswith("conv3x3_tc_kernel<")]
    assert len(row_pass) >= 10 and len(convs) >= 5            # d = 128 / 256 / 512 x operand-split variants; conv stage variants
    for k in row_pass + convs:
        c = sass[k]
        assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0 and c["UTCBAR"] > 0, (k, dict(c))
    for k in row_pass:
        assert sass[k]["STTM"] > 0, k                          # the transformed H tile goes back to tensor memory (TS-mode GEMM2)
    # hi + lo row planes double the MMAs of the single-plane kernel
    assert sass["drsa_tc_step_kernel<256, 0, 1>"]["UTCHMMA"] == 2 * sass["drsa_tc_step_kernel<256, 0, 0>"]["UTCHMMA"]


def test_no_legacy_tensor_core_path_anywhere(sass):
    for k, c in sass.items():
        assert c["HMMA"] == 0 and c["HGMMA"] == 0, k


def test_peer_exchange_instructions(sass):
    for k in ("finish_fused_kernel", "finish_small_kernel"):
        assert sass[k]["LDG.E.64.STRONG.SYS"] >= 8            # first look at up to 8 peers' words + the polling loop
        assert sass[k]["STG.E.128"] >= 7                       # the pushes (and other vector stores)
    assert sass["finish_fused_kernel"]["LDGSTS"] > 0           # operand panels staged by cp.async
