"""Long-horizon parity of the row-pass arithmetic modes against the on-device fp32 path (and, where a fixture exists,
against the reference's golden trajectory): 2 000 steps = the reference's default horizon (drsa.py:76).

  python scripts/horizon_parity.py [steps] [M,d ...]
"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import drsa_ref                                    # checker (synthetic rows + the angle metric)
from cxai.xai.drsa.drsa import SubspaceOptimizer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or \
    [(8192, 128), (65536, 128), (65536, 256), (262144, 256), (640000, 256)]
modes = os.environ.get("MODES", "fp32,tc,tc_split,tc_hilo,tc_dc,tc32").split(",")
K = 4
for M, d in shapes:
    A, C = drsa_ref.synth_pairs(M, d, seed=1000 + d + M % 997)
    U0 = drsa_ref.synth_U0(d, d, seed=7)
    res = {}
    for prec in modes:
        opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device="cuda", precision=prec)
        torch.cuda.synchronize(); t0 = time.time()
        opt.run(steps=steps, save=False)
        torch.cuda.synchronize(); dt = time.time() - t0
        res[prec] = (opt.obj_history.copy(), opt.U.cpu().clone(), dt)
        del opt
    o32, U32, _ = res[modes[0]]
    for prec in modes:
        o, U, dt = res[prec]
        rel = np.max(np.abs(o - o32) / np.abs(o32))
        ang = drsa_ref.principal_angle(U, U32, K)
        print(f"M={M} d={d} steps={steps} {prec:9s}: {dt / steps * 1e3:7.3f} ms/step  rel obj vs {modes[0]} {rel:.2e}  "
              f"angle {ang:.2e} rad  obj_end {o[-1]:.6f}", flush=True)
