"""Which fp16 rounding of the tensor-core row pass moves the trajectory?  (CPU simulation, see sim_precision.py)"""
import sys, os, math
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import drsa_ref
M, d, steps, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 4
torch.set_num_threads(8)
p2 = lambda a: 2.0 ** (7 - math.floor(math.log2(a)))
f16 = lambda x: x.to(torch.float16).to(torch.float64)
idt = lambda x: x


def run(A, C, U0, rA, rU, rP):
    A64, C64 = A.double(), C.double()
    sA, sC = p2(float(A.abs().max())), p2(float(C.abs().max()))
    A16, C16 = rA(A64 * sA), rA(C64 * sC)
    rhoA = float(A16.norm(dim=1).max()); rhoC = float(C16.norm(dim=1).max())
    pq = 2.0 ** math.floor(math.log2(32768.0 / max(rhoA * rhoC * rhoC, rhoA * rhoA * rhoC)))
    inv = 1.0 / (sA * sC)
    U = U0.double(); d_k = U.shape[1] // K; objs = []
    def evaluate(U):
        Uhi = rU(U)
        HA, HC = A16 @ Uhi, C16 @ Uhi
        s = (HA * HC).view(-1, K, d_k).sum(-1); g = torch.relu(s)
        sumsq = ((g * inv) ** 2).sum(0)
        gg = (g * pq).repeat_interleave(d_k, dim=1)
        P, Q = rP(gg * HC), rP(gg * HA)
        X = (A16.T @ P + C16.T @ Q) * (inv * inv / pq)
        obj, grad = drsa_ref.finish_from_sums(X.float().double(), sumsq.float().double(), A.shape[0], K)
        return obj + float((grad * (U - Uhi)).sum()), grad
    for _ in range(steps):
        obj, grad = evaluate(U); objs.append(float(obj))
        U = drsa_ref.orthogonalize((U + grad).float()).double()
    objs.append(float(evaluate(U)[0]))
    return np.asarray(objs), U

A, C = drsa_ref.synth_pairs(M, d, 77, structured=True)
U0 = drsa_ref.synth_U0(d, seed=78)
objs64, U64 = drsa_ref.run_closed_form(A, C, U0, K, steps)
for name, (rA, rU, rP) in {"none": (idt, idt, idt), "A,C": (f16, idt, idt), "U": (idt, f16, idt), "P,Q": (idt, idt, f16), "all": (f16, f16, f16)}.items():
    objs, U = run(A, C, U0, rA, rU, rP)
    print(f"rounded {name:5s}: obj err {np.max(np.abs(objs - objs64) / objs64):.2e}  angle {drsa_ref.principal_angle(U, U64, K):.2e}", flush=True)
