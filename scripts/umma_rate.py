"""tcgen05.mma issue rate for the shapes of the DRSA row pass (operands resident, no TMA, no epilogue)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drsa_audio_b200 import _lib as L
torch.cuda.init()
names = ["cta_group::1 SS M128 N128 (GEMM1, single-CTA kernel)", "cta_group::1 TS M128 N256 (GEMM2, single-CTA kernel)",
         "cta_group::2 SS M256 N128 (GEMM1, pair kernel)", "cta_group::2 TS M256 N256 (GEMM2, pair kernel)",
         "cta_group::2 SS M256 N256", "cta_group::1 SS M128 N256"]
for mode, name in enumerate(names):
    out = ctypes.c_float(0)
    st = L.lib().drsa_selftest_umma(10 + mode, ctypes.byref(out))
    print(f"{name:58s} status {st}  {out.value:7.1f} cycles per MMA (K = 16)")
