"""Device-side fp64 simulation of the tensor-core arithmetic modes over the reference's horizon: which fp16 rounding moves the
trajectory, and does error feedback on the per-step rounding of U (sigma-delta: the residual of one step is added before
the next rounding) remove its contribution?   python scripts/sim_feedback_gpu.py [steps] [M,d ...]"""
import os, sys, math
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import drsa_ref

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(8192, 128), (65536, 128), (65536, 256), (640000, 256)]
K = 4
dev = "cuda"
f16 = lambda x: x.to(torch.float16).to(torch.float64)
p2 = lambda a: 2.0 ** (7 - math.floor(math.log2(a)))


def polar(Y):
    S, V = torch.linalg.eigh(Y.T @ Y)
    return Y @ (V @ torch.diag(S.rsqrt()) @ V.T)


def run_dc(A, C, U0, every):
    """rows in fp16 + U in fp16 with error feedback + DEFERRED CORRECTION of the row rounding: every `every` steps the sums
    are evaluated with the full-precision rows as well and the difference is added to the cheap sums of the next steps."""
    sA, sC = p2(float(A.abs().max())), p2(float(C.abs().max()))
    Af, Cf = A * sA, C * sC
    Ah, Ch = f16(Af), f16(Cf)
    rhoA, rhoC = float(Ah.norm(dim=1).max()), float(Ch.norm(dim=1).max())
    pq = 2.0 ** math.floor(math.log2(32768.0 / max(rhoA * rhoC * rhoC, rhoA * rhoA * rhoC)))
    inv = 1.0 / (sA * sC)
    U = U0.clone(); d_k = U.shape[1] // K; M = A.shape[0]
    E = torch.zeros_like(U)

    def sums(As, Cs, Uh):
        HA, HC = As @ Uh, Cs @ Uh
        s = (HA * HC).view(-1, K, d_k).sum(-1); g = torch.relu(s)
        ss = ((g * inv) ** 2).sum(0)
        gg = (g * pq).repeat_interleave(d_k, dim=1)
        return (As.T @ f16(gg * HC) + Cs.T @ f16(gg * HA)) * (inv * inv / pq), ss
    objs = []
    dX = dss = 0.0
    for it in range(steps + 1):
        Uh = f16(U + E); E = U + E - Uh
        X, ss = sums(Ah, Ch, Uh)
        if it % every == 0:
            Xf, ssf = sums(Af, Cf, Uh)
            dX, dss = Xf - X, ssf - ss
        X, ss = X + dX, ss + dss
        q = torch.sqrt(ss / M); obj = torch.mean(torch.sqrt(q)) ** 2
        grad = X * (torch.sqrt(obj) / (K * M * q ** 1.5)).repeat_interleave(d_k)[None, :]
        objs.append(float(obj + (grad * (U - Uh)).sum()))
        if it < steps:
            U = polar(U + grad)
    return np.asarray(objs), U


def run(A, C, U0, round_rows, round_u, feedback, round_pq=True):
    sA, sC = p2(float(A.abs().max())), p2(float(C.abs().max()))
    As, Cs = A * sA, C * sC
    if round_rows:
        As, Cs = f16(As), f16(Cs)
    rhoA, rhoC = float(As.norm(dim=1).max()), float(Cs.norm(dim=1).max())
    pq = 2.0 ** math.floor(math.log2(32768.0 / max(rhoA * rhoC * rhoC, rhoA * rhoA * rhoC)))
    inv = 1.0 / (sA * sC)
    U = U0.clone(); d_k = U.shape[1] // K; M = A.shape[0]
    E = torch.zeros_like(U)
    objs = []
    for it in range(steps + 1):
        if round_u:
            Uh = f16(U + E) if feedback else f16(U)
            if feedback:
                E = U + E - Uh
        else:
            Uh = U
        HA, HC = As @ Uh, Cs @ Uh
        s = (HA * HC).view(-1, K, d_k).sum(-1); g = torch.relu(s)
        sumsq = ((g * inv) ** 2).sum(0)
        gg = (g * pq).repeat_interleave(d_k, dim=1)
        P, Q = gg * HC, gg * HA
        if round_pq:
            P, Q = f16(P), f16(Q)
        X = (As.T @ P + Cs.T @ Q) * (inv * inv / pq)
        q = torch.sqrt(sumsq / M); obj = torch.mean(torch.sqrt(q)) ** 2
        grad = X * (torch.sqrt(obj) / (K * M * q ** 1.5)).repeat_interleave(d_k)[None, :]
        objs.append(float(obj + (grad * (U - Uh)).sum()))
        if it < steps:
            U = polar(U + grad)
    return np.asarray(objs), U


for M, d in shapes:
    A, C = drsa_ref.synth_pairs(M, d, seed=1000 + d + M % 997)
    U0 = drsa_ref.synth_U0(d, d, seed=7)
    A, C, U0 = A.double().to(dev), C.double().to(dev), U0.double().to(dev)
    o0, Ut = run(A, C, U0, False, False, False, round_pq=False)
    for every in [int(v) for v in os.environ.get("DC", "").split(",") if v]:
        o, U = run_dc(A, C, U0, every)
        print(f"M={M} d={d} deferred correction every {every:3d}: rel obj {np.max(np.abs(o - o0) / o0):.2e}  angle "
              f"{drsa_ref.principal_angle(U.cpu(), Ut.cpu(), K):.2e}", flush=True)
    if os.environ.get("DC_ONLY"):
        continue
    for name, args in {"P,Q only": (False, False, False), "rows16": (True, False, False), "U16": (False, True, False),
                       "U16+feedback": (False, True, True), "rows16+U16 (tc)": (True, True, False),
                       "rows16+U16+feedback": (True, True, True)}.items():
        o, U = run(A, C, U0, *args)
        print(f"M={M} d={d} {name:22s}: rel obj {np.max(np.abs(o - o0) / o0):.2e}  angle {drsa_ref.principal_angle(U.cpu(), Ut.cpu(), K):.2e}",
              flush=True)
