"""Where the end-to-end time of SubspaceOptimizer(host rows).run() goes (cfg 2): H2D, packing, graph capture, steps, D2H."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_rows_cuda          # noqa: E402
from cxai.xai.drsa.drsa import SubspaceOptimizer          # noqa: E402
from bench import synth_U0          # noqa: E402

PREC = sys.argv[1] if len(sys.argv) > 1 else "tc"
dev = torch.device("cuda", 0)
M, d, K = 640_000, 256, 4
A, C = synth_rows_cuda(M, d, 20262, dev)
U0 = synth_U0(d, seed=5)
Ah, Ch = A.cpu().pin_memory(), C.cpu().pin_memory()
del A, C
torch.cuda.synchronize()


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(3):
    t0 = t()
    Ad = Ah.to(dev); Cd = Ch.to(dev)
    t1 = t()
    opt = SubspaceOptimizer(U0, Ad, Cd, None, num_concepts=K, device=dev, precision=PREC)
    t2 = t()
    opt.run(steps=4, save=False)
    t3 = t()
    opt.run(steps=500, save=False)
    t4 = t()
    Uh = opt.U.cpu()
    t5 = t()
    print(f"rep {rep}: H2D {1e3 * (t1 - t0):.1f} ms ({2 * M * d * 4 / (t1 - t0) / 1e9:.1f} GB/s)  construct(pack) {1e3 * (t2 - t1):.1f} ms  "
          f"first run(4) incl. capture {1e3 * (t3 - t2):.1f} ms  run(500) {1e3 * (t4 - t3):.1f} ms  D2H {1e3 * (t5 - t4):.2f} ms", flush=True)
    del opt, Ad, Cd
    t0 = t()
    opt = SubspaceOptimizer(U0, Ah, Ch, None, num_concepts=K, device=dev, precision=PREC)
    t1 = t()
    opt.run(steps=500, save=False)
    t2 = t()
    opt.run(steps=2000, save=False)
    t3 = t()
    print(f"        host rows: construct {1e3 * (t1 - t0):.1f} ms  run(500) {1e3 * (t2 - t1):.1f} ms  run(2000) {1e3 * (t3 - t2):.1f} ms", flush=True)
    del opt
