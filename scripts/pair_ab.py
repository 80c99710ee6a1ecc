"""A/B of the row pass at cfg 2: single-CTA kernel (variant 0) vs CTA-pair kernel (variant 2), CUDA events."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drsa_audio_b200 import _lib as L
from cxai.xai.drsa.drsa import SubspaceOptimizer
from bench import synth_rows_cuda
from bench import synth_U0
dev = torch.device("cuda", 0)
M, d, K = (int(sys.argv[1]) if len(sys.argv) > 1 else 640000), 256, 4
A, C = synth_rows_cuda(M, d, 20262, dev)
U0 = synth_U0(d, seed=5)
opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device=dev, precision="tc", use_cuda_graph=False)
opt._rows.split_u(opt._Uw)
res = {}
for rep in range(2):
    for v in (0, 2, 3):
        L.lib().drsa_debug_set_tc_variant(v)
        for _ in range(3): opt._rows.step(opt._Uw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): opt._rows.step(opt._Uw)
        e1.record(); torch.cuda.synchronize()
        res[v] = opt._rows.sums.clone()
        print(f"variant {v}: row pass {e0.elapsed_time(e1) / 100:.4f} ms  ({8.0 * M * d * d / (e0.elapsed_time(e1) / 100 * 1e-3) / 1e12:.0f} TFLOP/s)", flush=True)
L.lib().drsa_debug_set_tc_variant(0)
print("max rel diff of the sums:", float((res[0] - res[2]).abs().max() / res[0].abs().max()))
import ctypes
out = (ctypes.c_int * 5)()
L.lib().drsa_debug_tc_kernel_attrs(256, 2, out)
print("pair kernel: regs", out[0], "local bytes", out[3], "max active 2-CTA clusters", out[4])
