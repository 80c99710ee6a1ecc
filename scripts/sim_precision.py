"""CPU simulation of the arithmetic of the tensor-core row pass (which operands are rounded to fp16 and where) against
the reference trajectory.  Used to decide the precision mode; no GPU needed.

  python scripts/sim_precision.py [M] [d] [steps]
"""
import sys, os, math
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import drsa_ref

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
K = 4
torch.set_num_threads(8)


def p2(absmax):
    return 2.0 ** (7 - math.floor(math.log2(absmax)))


def f16(x):
    return x.to(torch.float16).to(torch.float64)


def run(A, C, U0, mode, correct):
    A64, C64 = A.double(), C.double()
    sA, sC = p2(float(A.abs().max())), p2(float(C.abs().max()))
    A16, C16 = f16(A64 * sA), f16(C64 * sC)
    rhoA = float(A16.norm(dim=1).max()); rhoC = float(C16.norm(dim=1).max())
    pq = 2.0 ** math.floor(math.log2(32768.0 / max(rhoA * rhoC * rhoC, rhoA * rhoA * rhoC)))
    inv = 1.0 / (sA * sC)
    U = U0.double()
    d_k = U.shape[1] // K
    objs = []
    def evaluate(U):
        Uhi = f16(U)
        Ulo = f16(U - Uhi)
        Ue = Uhi + Ulo if mode == "hilo" else Uhi
        HA, HC = A16 @ Ue, C16 @ Ue
        s = (HA * HC).view(-1, K, d_k).sum(-1)
        g = torch.relu(s)
        sumsq = ((g * inv) ** 2).sum(0)
        gg = (g * pq).repeat_interleave(d_k, dim=1)
        P, Q = f16(gg * HC), f16(gg * HA)
        X = (A16.T @ P + C16.T @ Q) * (inv * inv / pq)
        obj, grad = drsa_ref.finish_from_sums(X.float().double(), sumsq.float().double(), A.shape[0], K)
        if correct and mode == "hi":
            obj = obj + float((grad * (U - Uhi)).sum())
        return obj, grad
    for _ in range(steps):
        obj, grad = evaluate(U)
        objs.append(float(obj))
        U = drsa_ref.orthogonalize((U + grad).float()).double()
    objs.append(float(evaluate(U)[0]))
    return np.asarray(objs), U


for structured in (True, False):
    A, C = drsa_ref.synth_pairs(M, d, 77, structured=structured)
    U0 = drsa_ref.synth_U0(d, seed=78)
    objs_ref, U_ref = drsa_ref.run_autograd(A, C, U0, K, steps)
    objs64, U64 = drsa_ref.run_closed_form(A, C, U0, K, steps)
    print(f"structured={structured} M={M} d={d} steps={steps}: fp32 reference vs fp64 closed form: obj "
          f"{np.max(np.abs(objs_ref - objs64) / objs64):.2e} angle {drsa_ref.principal_angle(U_ref, U64, K):.2e}")
    for mode, correct in (("hilo", False), ("hi", False), ("hi", True)):
        objs, U = run(A, C, U0, mode, correct)
        rel = np.max(np.abs(objs - objs_ref) / np.abs(objs_ref))
        rel64 = np.max(np.abs(objs - objs64) / np.abs(objs64))
        print(f"  mode={mode:5s} corrected={correct!s:5s}: max rel obj err vs fp32 ref {rel:.2e} (vs fp64 {rel64:.2e}), "
              f"final angle vs fp32 ref {drsa_ref.principal_angle(U, U_ref, K):.2e} rad (vs fp64 {drsa_ref.principal_angle(U, U64, K):.2e})")
