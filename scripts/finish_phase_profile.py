"""%globaltimer stamps of CTA 0 at the phase boundaries of the fused finish kernel (cfg2 shape by default)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drsa_audio_b200 import _lib as L
from cxai.xai.drsa.drsa import SubspaceOptimizer
from bench import synth_rows_cuda
d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M, K = 64000, 4
dev = torch.device("cuda")
A, C = synth_rows_cuda(M, d, 1, dev)
U0 = torch.linalg.qr(torch.randn(d, d))[0]
opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, precision="tc", use_cuda_graph=False)
opt._rows.split_u(opt._Uw)
opt.reset_log(64)
for _ in range(3): opt._step(opt._obj_log, -1, True)
buf = torch.zeros(64, dtype=torch.int64, device=dev)
L.lib().drsa_debug_set_tc_profile(buf.data_ptr())
opt._rows.step(opt._Uw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); opt._rows.finish(opt._Uw, opt.M_global, opt._obj_log, -1, True, opt.retraction_iters, opt.retraction_tol); e1.record()
torch.cuda.synchronize()
L.lib().drsa_debug_set_tc_profile(None)
v = buf.cpu().tolist(); n = v[15]; st = v[16:16 + n]
print(f"finish {e0.elapsed_time(e1)*1e3:.1f} us (events), {n} stamps, first->last {(st[-1]-st[0])/1e3:.1f} us")
print("deltas (us):", " ".join(f"{(b-a)/1e3:.1f}" for a, b in zip(st[:-1], st[1:])))
# sweeps of the polar iteration and step time along a 2 000-step run (the first steps take large ascent steps)
if len(sys.argv) > 2:
    A, C = synth_rows_cuda(640000, d, 1, dev)
    opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, precision=sys.argv[2])
    opt._rows.split_u(opt._Uw)
    opt.reset_log(2100)
    done = 0
    for upto in (8, 25, 50, 100, 200, 400, 800, 1200, 1600, 2000):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); opt.enqueue_steps(upto - done); e1.record(); torch.cuda.synchronize()
        print(f"steps {done:4d}..{upto:4d}: {e0.elapsed_time(e1) / (upto - done):.4f} ms/step, sweeps of the last step "
              f"{int(opt._rows.status[0])}", flush=True)
        done = upto
