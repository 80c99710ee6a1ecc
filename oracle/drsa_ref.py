"""CPU restatement of the reference's DRSA stage (stage 2 of the hot path).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every function follows ``/root/reference/cxai/xai/drsa/drsa.py`` at the cited
lines.  Parity status: PINNED -- ``oracle/gen_golden.py`` imports the unmodified
reference file in the build container (through a ``pathilib -> pathlib`` alias, the
file has a typo at drsa.py:4) and stores its outputs under ``tests/golden/``;
``tests/test_oracle_drsa.py`` checks this restatement against those vectors.

Two flavours are provided:
  * ``*_autograd``: the same torch ops in the same order as the reference, fp32,
    gradient by autograd (what the reference executes).
  * closed form: objective, per-concept sums of squares and the
    un-normalised gradient blocks X_k, in any dtype (fp64 for tolerance studies).
    This is the decomposition the CUDA kernels implement.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- objective
def generalized_fmean(x: torch.Tensor, p: float = 0.5) -> torch.Tensor:
    """F-mean with F(t)=t^p over dim 0 -- drsa.py:171-182."""
    return torch.pow(torch.mean(torch.pow(x, p), dim=0), 1 / p)


def objective_fn(r: torch.Tensor) -> torch.Tensor:
    """p=2 mean over rows, then p=0.5 mean over concepts -- drsa.py:224-238."""
    return generalized_fmean(generalized_fmean(r, 2), 0.5)


def obj_val(act, ctx, U, num_concepts: int, d_k: int) -> torch.Tensor:
    """drsa.py:123-155: project, multiply, block-sum, ReLU, pool."""
    xa = torch.matmul(act, U)
    xc = torch.matmul(ctx, U)
    x = torch.mul(xa, xc).view(-1, num_concepts, d_k)
    return objective_fn(F.relu(torch.sum(x, dim=-1)))


def orthogonalize(U: torch.Tensor) -> torch.Tensor:
    """Polar / Loewdin retraction U (U^T U)^(-1/2) -- drsa.py:201-221.

    Gram matrix in the dtype of U, eigen-decomposition in fp64 on the CPU, inverse
    square root and the final product formed in the dtype of U (fp32 in the
    reference)."""
    UtU = torch.matmul(U.T, U)
    S, V = torch.linalg.eigh(UtU.cpu().double())
    V = V.to(U.dtype)
    inv = torch.matmul(torch.matmul(V, torch.diag(1.0 / torch.sqrt(S.to(U.dtype)))), V.T).to(U)
    return torch.matmul(U, inv)


def orthogonalize_qr(Y: torch.Tensor) -> torch.Tensor:
    """Q of the thin QR factorisation with diag(R) > 0 (Householder QR + sign fix).  NOT the reference's retraction; the
    yardstick of the non-default ``retraction='qr'`` option only."""
    Q, R = torch.linalg.qr(Y.double())
    return (Q * torch.sign(torch.diagonal(R))[None, :]).to(Y.dtype)


def step_autograd(act, ctx, U, num_concepts: int):
    """One iteration of drsa.py:84-104: objective, autograd gradient, U <- polar(U + grad)."""
    d_k = U.shape[1] // num_concepts
    U = U.detach().clone().requires_grad_(True)
    obj = obj_val(act, ctx, U, num_concepts, d_k)
    obj.backward()
    with torch.no_grad():
        U_new = orthogonalize(U + U.grad)
    return obj.detach(), U.grad.detach(), U_new.detach()


def run_autograd(act, ctx, U0, num_concepts: int, steps: int):
    """drsa.py:76-117: `steps` updates, objective logged before each update plus one
    final evaluation -> steps+1 values."""
    U = U0.detach().clone()
    objs = []
    d_k = U.shape[1] // num_concepts
    for _ in range(steps):
        o, _, U = step_autograd(act, ctx, U, num_concepts)
        objs.append(float(o))
    objs.append(float(obj_val(act, ctx, U, num_concepts, d_k)))
    return np.asarray(objs, dtype=np.float64), U


# --------------------------------------------------------------------------- closed form
def step_sums(act, ctx, U, num_concepts: int):
    """The two row-sums one pass over (A, C) has to produce:

        sumsq[k] = sum_m relu(s_mk)^2
        X[:, block k] = A^T (r_k * HC_k) + C^T (r_k * HA_k)        (un-normalised)

    Both are plain sums over rows, so they add across row shards / ranks."""
    M = act.shape[0]
    m = U.shape[1]
    d_k = m // num_concepts
    HA = act @ U
    HC = ctx @ U
    s = (HA * HC).view(M, num_concepts, d_k).sum(-1)
    r = torch.relu(s)
    sumsq = (r * r).sum(0)
    rr = r.repeat_interleave(d_k, dim=1)
    X = act.T @ (rr * HC) + ctx.T @ (rr * HA)
    return X, sumsq


def finish_from_sums(X, sumsq, M_global: int, num_concepts: int):
    """objective and gradient from the (all-reduced) sums.

        q_k = sqrt(sumsq_k / M);  obj = (mean_k sqrt(q_k))^2
        grad[:, block k] = sqrt(obj) / (K M q_k^1.5) * X_k
    """
    K = num_concepts
    d_k = X.shape[1] // K
    q = torch.sqrt(sumsq / M_global)
    obj = torch.mean(torch.sqrt(q)) ** 2
    coef = torch.sqrt(obj) / (K * M_global * q ** 1.5)
    grad = X * coef.repeat_interleave(d_k)[None, :]
    return obj, grad


def step_closed_form(act, ctx, U, num_concepts: int):
    X, sumsq = step_sums(act, ctx, U, num_concepts)
    obj, grad = finish_from_sums(X, sumsq, act.shape[0], num_concepts)
    return obj, grad, orthogonalize(U + grad)


def run_closed_form(act, ctx, U0, num_concepts: int, steps: int, dtype=torch.float64):
    A, C, U = act.to(dtype), ctx.to(dtype), U0.to(dtype)
    objs = []
    for _ in range(steps):
        o, _, U = step_closed_form(A, C, U, num_concepts)
        objs.append(float(o))
    X, sumsq = step_sums(A, C, U, num_concepts)
    objs.append(float(finish_from_sums(X, sumsq, A.shape[0], num_concepts)[0]))
    return np.asarray(objs), U


def subspace_relevances(act, ctx, U, n_concepts: int = 4):
    """Per-instance R_k summed over positions, no ReLU -- explainer.py:206-242.
    act/ctx: [B, P, d] -> [B, K]."""
    if act.dim() == 2:
        act, ctx = act[None], ctx[None]
    B = act.shape[0]
    d_c = U.shape[0] // n_concepts
    x = (act @ U) * (ctx @ U)
    x = x.transpose(-2, -1).contiguous().view(B, n_concepts, -1, d_c)
    return x.sum(-1).sum(-1)


# --------------------------------------------------------------------------- preprocessing bits
def compute_context_vectors(a, R):
    """c = R / (a + 1e-7) -- preprocessing.py:179-193."""
    return R / (a + 1e-7)


def normalize_vectors(v):
    """v / sqrt(mean(v^2)) / d^0.25, statistic over ALL elements -- preprocessing.py:219-231."""
    d = v.shape[-1]
    E = torch.sqrt(torch.mean(torch.square(v)))
    return v / E / d ** 0.25


def vectors_from_maps_all(maps):
    """[N, d, H, W] -> [N*H*W, d], all positions (corrected layout)."""
    N, d = maps.shape[:2]
    return maps.reshape(N, d, -1).transpose(1, 2).reshape(-1, d)


def vectors_from_maps_ref(maps, idcs):
    """Bit-compatible restatement of preprocessing.py:234-256 (including its row
    scrambling): advanced index -> [B, L, d], then the reference's extra
    transpose(-2,-1).reshape(-1, d)."""
    B, d = maps.shape[:2]
    flat = maps.reshape(B, d, -1)
    v = flat[np.arange(B)[:, None], :, idcs]
    return v.transpose(-2, -1).reshape(-1, d)


def vectors_from_maps_fixed(maps, idcs):
    """Corrected gather: row (b, l) is the channel vector at position idcs[b, l]."""
    B, d = maps.shape[:2]
    flat = maps.reshape(B, d, -1)
    v = flat[np.arange(B)[:, None], :, idcs]          # [B, L, d]
    return v.reshape(-1, d)


# --------------------------------------------------------------------------- metrics
def principal_angle(U1, U2, num_concepts: int) -> float:
    """Largest principal angle (rad) between matching concept subspaces, sine
    formulation in fp64 (acos has a 1e-3 rad noise floor in fp32)."""
    U1 = torch.as_tensor(U1).double()
    U2 = torch.as_tensor(U2).double()
    d, m = U1.shape
    d_k = m // num_concepts
    worst = 0.0
    eye = torch.eye(d, dtype=torch.float64)
    for k in range(num_concepts):
        Q1 = torch.linalg.qr(U1[:, k * d_k:(k + 1) * d_k])[0]
        Q2 = torch.linalg.qr(U2[:, k * d_k:(k + 1) * d_k])[0]
        s = torch.linalg.svdvals((eye - Q1 @ Q1.T) @ Q2)
        worst = max(worst, float(torch.asin(s.clamp(max=1.0)).max()))
    return worst


# --------------------------------------------------------------------------- synthetic data
def synth_pairs(M: int, d: int, seed: int, structured: bool = True):
    """Synthetic (A, C) pairs of SURVEY section 8(d): A = relu(randn)*mask,
    C = randn*(A>0) (c = 0 wherever a = 0), optionally with a linear dependence of C
    on A so that concepts exist; both passed through normalize_vectors."""
    g = torch.Generator().manual_seed(seed)
    A = torch.relu(torch.randn(M, d, generator=g)) * (torch.rand(M, d, generator=g) < 0.7)
    C = torch.randn(M, d, generator=g) * (A > 0)
    if structured:
        W = torch.randn(d, d, generator=g) / d ** 0.5
        C = C + 0.5 * (A @ W) * (A > 0)
    return normalize_vectors(A).contiguous(), normalize_vectors(C).contiguous()


def synth_U0(d: int, m: int | None = None, seed: int = 5):
    g = torch.Generator().manual_seed(seed)
    Q = torch.linalg.qr(torch.randn(d, d, generator=g))[0]
    return Q[:, : (m or d)].contiguous()
