"""Per-phase cycle counters of the tensor-core row pass (CTA 0), cfg2 shape."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drsa_audio_b200 import _lib as L
from cxai.xai.drsa.drsa import SubspaceOptimizer
from bench import synth_rows_cuda
M, d, K = 640000, 256, 4
dev = torch.device("cuda")
A, C = synth_rows_cuda(M, d, 1, dev)
U0 = torch.linalg.qr(torch.randn(d, d))[0]
opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, precision=(sys.argv[1] if len(sys.argv) > 1 else "tc"), use_cuda_graph=False)
opt._rows.split_u(opt._Uw)
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L.lib().drsa_debug_set_tc_variant(variant)
for _ in range(3): opt._rows.step(opt._Uw)
buf = torch.zeros(16, dtype=torch.int64, device=dev)
L.lib().drsa_debug_set_tc_profile(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); opt._rows.step(opt._Uw); e1.record(); torch.cuda.synchronize()
L.lib().drsa_debug_set_tc_profile(None)
L.lib().drsa_debug_set_tc_variant(0)
sub = 32 if variant == 1 else 64; tiles = -(-M // sub); nrb = 148 // 2; per = -(-tiles // nrb)
v = buf.cpu().tolist()
print(f"row pass {e0.elapsed_time(e1):.3f} ms, {sub}-row subtiles per CTA ~{per}")
names = ["mma: GEMM1 issue", "mma: GEMM2 wait for first P chunk", "mma: GEMM2 issue", "epi: wait GEMM1", "epi: work"]
if variant == 2:
    names += ["mma: of which waiting for TMA (full barriers)", "epi (peer CTA): wait GEMM1", "epi (peer CTA): work", "-",
              "epi: tcgen05.ld + wait", "epi: products + transpose-reduce", "epi: bar.sync of the chunk's 4 warps",
              "epi: g, P/Q (smem reads, cvt)", "epi: tcgen05.st + wait::st", "epi: fence, syncwarp, arrive on the leader", "tma (peer): wait for a free stage", "tma (leader): wait for a free stage"]
    print("  (epilogue counters are per epilogue set: each set handles every other subtile)")
for name, x in zip(names, v):
    print(f"  {name:42s} {x:10d} cycles total, {x/per:9.0f} per subtile")
