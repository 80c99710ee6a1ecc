"""GPU parity tests of the DRSA stage (run on the B200 box: pytest -m gpu).  Everything goes
through the C ABI (ctypes) / the cxai mirror; the CPU oracle is the checker.

Tolerances (BASELINE.json north_star): relative error <= 1e-4 on the objective at every step,
principal angles between final subspaces <= 1e-3 rad."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import drsa_ref

pytestmark = pytest.mark.gpu

OBJ_TOL = 1e-4
ANGLE_TOL = 1e-3


@pytest.fixture(scope="module")
def L():
    from drsa_audio_b200 import _lib
    return _lib


def _dev(t):
    return t.cuda().contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _sums_gpu(L, A, C, U, K, prec):
    lib = L.lib()
    M, d = A.shape
    m = U.shape[1]
    code = {"tc": L.PREC_TC_F16, "tc_split": L.PREC_TC_F16X2, "fp32": L.PREC_FP32}[prec]
    ws = torch.empty(int(L.check(lib.drsa_step_workspace_bytes(M, d, m, K, code))), dtype=torch.uint8, device="cuda")
    sums = torch.full((d * m + K,), float("nan"), device="cuda")
    Ad, Cd, Ud = _dev(A), _dev(C), _dev(U)
    s = torch.cuda.current_stream().cuda_stream
    if prec != "fp32":
        import math
        from cxai.xai.drsa.drsa import _pow2_scale
        sA, sC = _pow2_scale(float(A.abs().max())), _pow2_scale(float(C.abs().max()))
        rA, rC = float(A.norm(dim=1).max()) * sA, float(C.norm(dim=1).max()) * sC
        pq = 2.0 ** math.floor(math.log2(32768.0 / max(rA * rC * rC, rA * rA * rC)))
        A16 = torch.empty(M, d, dtype=torch.float16, device="cuda")
        C16 = torch.empty(M, d, dtype=torch.float16, device="cuda")
        L.check(lib.drsa_pack_f16(_ptr(Ad), Ad.numel(), sA, _ptr(A16), s))
        L.check(lib.drsa_pack_f16(_ptr(Cd), Cd.numel(), sC, _ptr(C16), s))
        hi = torch.empty(m, d, dtype=torch.float16, device="cuda")
        lo = torch.empty(m, d, dtype=torch.float16, device="cuda") if prec == "tc_split" else None
        L.check(lib.drsa_split_u(_ptr(Ud), d, m, _ptr(hi), _ptr(lo), s))
        L.check(lib.drsa_step(_ptr(A16), _ptr(C16), None, _ptr(hi), _ptr(lo), M, d, m, K, code, sA, sC, pq,
                              _ptr(sums), _ptr(ws), ws.numel(), s), "drsa_step tc")
    else:
        L.check(lib.drsa_step(_ptr(Ad), _ptr(Cd), _ptr(Ud), None, None, M, d, m, K, code, 1.0, 1.0, 1.0,
                              _ptr(sums), _ptr(ws), ws.numel(), s), "drsa_step fp32")
    torch.cuda.synchronize()
    out = sums.cpu().double()
    return out[: d * m].view(d, m), out[d * m:]


# ------------------------------------------------------------------ building blocks
@pytest.mark.parametrize("variant", [0, 1])
def test_umma_selftest(L, variant):
    """tcgen05 descriptors in isolation: 0 = SS K-major x K-major, 1 = TS (TMEM A) x MN-major B.
    Inputs are small dyadic rationals, so the fp32 result must be exact."""
    err = ctypes.c_float(-1.0)
    L.check(L.lib().drsa_selftest_umma(variant, ctypes.byref(err)), "selftest")
    assert err.value == 0.0, err.value


@pytest.mark.parametrize("M,d,m,K", [(333, 32, 32, 2), (1000, 64, 64, 4), (777, 64, 32, 2), (4096, 128, 128, 4),
                                     (2500, 256, 256, 4), (300, 96, 96, 3)])
def test_row_sums_fp32_match_oracle(L, M, d, m, K):
    A, C = drsa_ref.synth_pairs(M, d, 100 + d)
    U = drsa_ref.synth_U0(d, m, 7)
    X, ss = _sums_gpu(L, A, C, U, K, "fp32")
    Xr, ssr = drsa_ref.step_sums(A.double(), C.double(), U.double(), K)
    assert torch.linalg.norm(X - Xr) / torch.linalg.norm(Xr) < 2e-6
    np.testing.assert_allclose(ss.numpy(), ssr.numpy(), rtol=2e-6)


@pytest.mark.parametrize("prec", ["tc", "tc_split"])
@pytest.mark.parametrize("M,d,m,K", [(1, 128, 128, 4), (63, 128, 128, 4), (128, 128, 128, 4), (1000, 128, 128, 2), (5000, 128, 128, 1),
                                     (4096, 256, 256, 4), (20011, 256, 256, 8), (40000, 256, 256, 2), (3000, 256, 128, 2),
                                     (777, 512, 512, 8), (20000, 512, 512, 8), (9000, 512, 256, 4)])
def test_row_sums_tensor_core_match_oracle(L, M, d, m, K, prec):
    """fused tcgen05 kernel vs fp64 oracle; fp16 storage of the rows bounds the error (2^-12 per
    element, averaged over rows).  'tc' evaluates at fp16(U) by definition, 'tc_split' at hi + lo = U."""
    if d == 512 and prec == "tc_split":
        pytest.skip("U^T hi + lo of a 128-column group is 256 KB at d = 512: only the single-pass mode exists")
    A, C = drsa_ref.synth_pairs(max(M, 2), d, 200 + d + K)
    A, C = A[:M].contiguous(), C[:M].contiguous()
    U = drsa_ref.synth_U0(d, m, 9)
    X, ss = _sums_gpu(L, A, C, U, K, prec)
    # (1) kernel correctness: oracle on the SAME fp16-quantised operands (only P/Q rounding and summation order differ)
    Uq = U.half().double() if prec == "tc" else U.double()
    Xq, ssq = drsa_ref.step_sums(A.half().double(), C.half().double(), Uq, K)
    relq = float(torch.linalg.norm(X - Xq) / torch.linalg.norm(Xq))
    assert relq < (4e-4 if M < 4096 else 1.5e-4), relq      # fp16 rounding of P/Q: 2^-12 per element, averaged over rows
    # (a single row has no averaging: a concept whose s cancels to ~0 only matches in absolute terms)
    np.testing.assert_allclose(ss.numpy(), ssq.numpy(), rtol=2e-5, atol=(1e-3 if M < 64 else 0.0) * float(ssq.max()))
    # (2) arithmetic mode vs the exact oracle: fp16 storage costs 2^-12 per element, averaged over rows
    Xr, ssr = drsa_ref.step_sums(A.double(), C.double(), U.double(), K)
    relX = float(torch.linalg.norm(X - Xr) / torch.linalg.norm(Xr))
    loose = M < 4096 or prec == "tc"          # 'tc' adds the (systematic) 2^-12 rounding of U
    assert relX < (1e-3 if loose else 2e-4), relX
    np.testing.assert_allclose(ss.numpy(), ssr.numpy(), rtol=1e-3 if loose else 2e-4,
                               atol=(1e-2 if M < 64 else 0.0) * float(ssr.max()))


@pytest.mark.parametrize("M,d,m,K", [(31, 128, 128, 4), (5000, 128, 128, 2), (20011, 256, 256, 8), (3000, 256, 128, 2)])
def test_row_sums_tmem_resident_u_variant(L, M, d, m, K):
    """The experimental row-pass variant with U^T in tensor memory (drsa_debug_set_tc_variant(1)) computes the same sums."""
    A, C = drsa_ref.synth_pairs(M, d, 300 + d + K)
    U = drsa_ref.synth_U0(d, m, 9)
    L.lib().drsa_debug_set_tc_variant(1)
    try:
        X, ss = _sums_gpu(L, A, C, U, K, "tc")
    finally:
        L.lib().drsa_debug_set_tc_variant(0)
    X0, ss0 = _sums_gpu(L, A, C, U, K, "tc")
    Xq, ssq = drsa_ref.step_sums(A.half().double(), C.half().double(), U.half().double(), K)
    assert float(torch.linalg.norm(X - Xq) / torch.linalg.norm(Xq)) < (4e-4 if M < 4096 else 1.5e-4)
    np.testing.assert_allclose(ss.numpy(), ssq.numpy(), rtol=2e-5, atol=(1e-3 if M < 64 else 0.0) * float(ssq.max()))
    assert float(torch.linalg.norm(X - X0) / torch.linalg.norm(X0)) < 3e-4


@pytest.mark.parametrize("M,d,m,K", [(31, 256, 256, 4), (5000, 256, 256, 2), (20011, 256, 256, 8), (150000, 256, 256, 4)])
def test_row_sums_cta_pair_variant(L, M, d, m, K):
    """The CTA-pair row pass (cta_group::2: the two column groups of a row block share the staged rows,
    drsa_debug_set_tc_variant(2)) computes the same sums as the single-CTA kernel."""
    A, C = drsa_ref.synth_pairs(M, d, 400 + M % 97 + K)
    U = drsa_ref.synth_U0(d, m, 9)
    L.lib().drsa_debug_set_tc_variant(2)
    try:
        X, ss = _sums_gpu(L, A, C, U, K, "tc")
    finally:
        L.lib().drsa_debug_set_tc_variant(0)
    X0, ss0 = _sums_gpu(L, A, C, U, K, "tc")
    Xq, ssq = drsa_ref.step_sums(A.half().double(), C.half().double(), U.half().double(), K)
    assert float(torch.linalg.norm(X - Xq) / torch.linalg.norm(Xq)) < (4e-4 if M < 4096 else 1.5e-4)
    np.testing.assert_allclose(ss.numpy(), ssq.numpy(), rtol=2e-5, atol=(1e-3 if M < 64 else 0.0) * float(ssq.max()))
    assert float(torch.linalg.norm(X - X0) / torch.linalg.norm(X0)) < 3e-4


@pytest.mark.parametrize("d,K,M", [(100, 4, 30000), (64, 4, 30000), (96, 2, 20000)])
def test_zero_padded_problem_runs_on_tensor_cores_and_matches_reference(d, K, M):
    """Split layers that are not tensor-core shapes (arch A layer 19: d = 100, K = 4 x 25) run zero-padded (every concept
    block widened to 32 / 64 columns, unit columns in new all-zero channels): trajectory and final subspaces must match the
    reference algorithm like any other tensor-core shape, and U comes back in the original d x d layout, orthonormal."""
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    steps = 40
    A, C = drsa_ref.synth_pairs(M, d, 900 + d)
    U0 = drsa_ref.synth_U0(d, seed=901)
    opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device="cuda", precision="tc")
    assert opt._pad is not None and opt.precision == "tc" and tuple(opt.U.shape) == (d, d)
    assert torch.equal(opt.U.cpu(), U0)
    opt.run(steps=steps, save=False)
    objs_ref, U_ref = drsa_ref.run_autograd(A, C, U0, K, steps)
    rel = float(np.max(np.abs(opt.obj_history - objs_ref) / np.abs(objs_ref)))
    ang = drsa_ref.principal_angle(opt.U.cpu(), U_ref, K)
    assert rel < 1e-4 and ang < 1e-3, (rel, ang)
    U = opt.U
    assert float((U.T @ U - torch.eye(d, device=U.device)).abs().max()) < 1e-5
    # 'auto' keeps exact fp32 arithmetic for small problems (padded only where the fused finish kernel needs multiples of 32)
    small = SubspaceOptimizer(U0, A[:4096], C[:4096], None, num_concepts=K, device="cuda")
    assert small.precision == "fp32" and (small._pad is not None) == (d > 64 and d % 32 != 0)
    small.run(steps=5, save=False)
    objs5, U5 = drsa_ref.run_autograd(A[:4096], C[:4096], U0, K, 5)
    assert float(np.max(np.abs(small.obj_history - objs5) / np.abs(objs5))) < 1e-5
    assert drsa_ref.principal_angle(small.U.cpu(), U5, K) < 1e-4


def test_tensor_core_scale_invariance(L):
    """power-of-two pre-scaling of the fp16 rows is undone exactly: unnormalised inputs 1000x larger
    give row sums 1e6x / 1e12x larger."""
    M, d, K = 2048, 128, 4
    A, C = drsa_ref.synth_pairs(M, d, 5)
    U = drsa_ref.synth_U0(d, d, 9)
    X1, s1 = _sums_gpu(L, A, C, U, K, "tc")
    X2, s2 = _sums_gpu(L, A * 1024.0, C * 1024.0, U, K, "tc")
    np.testing.assert_allclose((X2 / 1024.0 ** 4).numpy(), X1.numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose((s2 / 1024.0 ** 4).numpy(), s1.numpy(), rtol=1e-6)


@pytest.mark.parametrize("d,m", [(32, 32), (64, 64), (64, 32), (128, 128), (256, 256), (512, 512)])
def test_polar_retraction_matches_reference_formula(d, m):
    from cxai.xai.drsa.drsa import orthogonalize
    g = torch.Generator().manual_seed(d + m)
    U = drsa_ref.synth_U0(d, m, 3)
    Y = U + 0.3 * torch.randn(d, m, generator=g) / d ** 0.5
    got = orthogonalize(Y.cuda()).cpu()
    want = drsa_ref.orthogonalize(Y.double())
    assert float((got.double() - want).abs().max()) < 5e-6
    G = got.double().T @ got.double()
    assert float((G - torch.eye(m, dtype=torch.float64)).abs().max()) < 5e-6


def test_polar_retraction_far_from_orthogonal():
    from cxai.xai.drsa.drsa import orthogonalize
    g = torch.Generator().manual_seed(1)
    Y = torch.randn(64, 64, generator=g)           # condition number ~ 1e2-1e3
    got = orthogonalize(Y.cuda(), max_iters=40).cpu().double()
    want = drsa_ref.orthogonalize(Y.double())
    assert float((got - want).abs().max()) < 1e-4


# ------------------------------------------------------------------ trajectories vs golden vectors of the reference
def _golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"drsa_{name}.npz"))
    M, d, K = int(g["M"]), int(g["d"]), int(g["K"])
    if "A" in g.files:
        A, C = torch.from_numpy(g["A"]), torch.from_numpy(g["C"])
    else:
        A, C = drsa_ref.synth_pairs(M, d, int(g["seed"]))
    U0 = torch.from_numpy(g["U0"]) if "U0" in g.files else drsa_ref.synth_U0(d, int(g["m"]), int(g["seed"]) + 1)
    return g, A, C, U0, K


@pytest.mark.parametrize("name,prec,graph", [("tiny", "fp32", False), ("ragged", "fp32", True), ("toy64", "fp32", True),
                                             ("d128", "fp32", False), ("d256", "fp32", True), ("rect", "fp32", True),
                                             ("d128", "tc", True), ("d256", "tc", True), ("d256", "tc", False),
                                             ("d128", "tc_split", True), ("d256", "tc_split", False),
                                             ("d512", "fp32", True), ("d512", "tc", True),
                                             ("d128", "tc_hilo", True), ("d256", "tc_hilo", False), ("d512", "tc_hilo", True),
                                             ("d128", "tc32", False), ("d256", "tc32", True),
                                             ("d128", "tc_dc", True), ("d256", "tc_dc", False), ("d512", "tc_dc", True)])
def test_run_matches_reference_golden(golden_dir, name, prec, graph, tmp_path):
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    g, A, C, U0, K = _golden(golden_dir, name)
    steps = int(g["steps"])
    opt = SubspaceOptimizer(U0, A, C, str(tmp_path), num_concepts=K, device="cuda", precision=prec,
                            use_cuda_graph=graph)
    opt.run(steps=steps)
    objs = opt.obj_history
    assert len(objs) == steps + 1
    rel = np.max(np.abs(objs - g["objs"]) / np.abs(g["objs"]))
    ang = drsa_ref.principal_angle(opt.U.cpu(), g["U_final"], K)
    print(f"{name}/{prec}: rel obj {rel:.2e} angle {ang:.2e} sweeps {opt.last_status}")
    assert rel < OBJ_TOL, rel
    assert ang < ANGLE_TOL, ang
    assert opt.last_status[1] == 0                      # retraction converged
    # on-disk formats of drsa.py:157-168
    import pickle
    with open(tmp_path / "projection_matrix.pkl", "rb") as f:
        Usaved = pickle.load(f)
    assert Usaved.dtype == np.float32 and Usaved.shape == tuple(U0.shape)
    lines = open(tmp_path / "train_stats.csv").read().splitlines()
    assert lines[0] == ",loss" and len(lines) == steps + 2
    UtU = opt.U.T @ opt.U
    assert float((UtU - torch.eye(UtU.shape[0], device="cuda")).abs().max()) < 5e-6


def test_main_matches_reference_result_tree(golden_dir, tmp_path):
    """drsa.main (drsa.py:241-300) against the result tree the reference's own main wrote for the same seed
    (oracle/gen_golden_main.py): numpy seed -> ortho_group.rvs -> compounding column permutations -> run{r}/ files."""
    import pickle
    from cxai.xai.drsa import drsa
    g = np.load(os.path.join(golden_dir, "drsa_main.npz"))
    M, d, K, steps, runs = (int(g[k]) for k in ("M", "d", "K", "steps", "runs"))
    A, C = drsa_ref.synth_pairs(M, d, int(g["row_seed"]))
    drsa.main(A, C, str(tmp_path), num_concepts=K, steps=steps, runs=runs, seed=int(g["seed"]), device="cuda")
    assert sorted(os.listdir(tmp_path)) == [f"run{r}" for r in range(1, runs + 1)]
    for r in range(1, runs + 1):
        lines = open(tmp_path / f"run{r}" / "train_stats.csv").read().splitlines()
        assert lines[0] == str(g["header"]) and len(lines) == steps + 2
        assert [l.split(",")[0] for l in lines[1:]] == [str(i) for i in range(steps + 1)]        # pandas index column
        loss = np.asarray([float(l.split(",")[1]) for l in lines[1:]])
        want = g[f"loss_run{r}"]
        rel = float(np.max(np.abs(loss - want) / np.abs(want)))
        with open(tmp_path / f"run{r}" / "projection_matrix.pkl", "rb") as f:
            U = pickle.load(f)
        assert isinstance(U, np.ndarray) and U.dtype == np.float32 and U.shape == (d, d)
        ang = drsa_ref.principal_angle(torch.from_numpy(U), torch.from_numpy(g[f"U_run{r}"]), K)
        print(f"main run{r}: rel obj {rel:.2e} angle {ang:.2e} max |dU| {np.abs(U - g[f'U_run{r}']).max():.2e}")
        assert rel < OBJ_TOL and ang < ANGLE_TOL
        assert np.abs(U - g[f"U_run{r}"]).max() < 1e-4          # same start (same permutation), same trajectory
    # the three runs start from different column orders: their logs differ
    assert abs(g["loss_run1"][0] - g["loss_run2"][0]) > 1e-6


def test_obj_val_static_and_autograd(golden_dir):
    from cxai.xai.drsa.drsa import SubspaceOptimizer, objective_fn
    g, A, C, U0, K = _golden(golden_dir, "toy64")
    U = U0.clone().cuda().requires_grad_(True)
    obj = SubspaceOptimizer.obj_val(A.cuda(), C.cuda(), U, objective_fn, K, U0.shape[1] // K)
    obj.backward()
    assert abs(float(obj) - float(g["obj0"])) / float(g["obj0"]) < 1e-5
    rel = np.linalg.norm(U.grad.cpu().numpy() - g["grad0"]) / np.linalg.norm(g["grad0"])
    assert rel < 1e-5, rel


def test_first_order_objective_correction(L):
    """DRSA_PREC_TC_F16 evaluates the sums at fp16(U); drsa_finish_step(u_rounded=1) must log
    f(fp16 U) + <grad, U - fp16 U>, which matches f(U) to second order in the rounding."""
    lib = L.lib()
    M, d, K = 30000, 128, 4
    A, C = drsa_ref.synth_pairs(M, d, 31)
    U = drsa_ref.synth_U0(d, d, 32)
    Uq = U.half().double()
    Xq, ssq = drsa_ref.step_sums(A.double(), C.double(), Uq, K)
    f_q, grad_q = drsa_ref.finish_from_sums(Xq, ssq, M, K)
    f_true = float(drsa_ref.finish_from_sums(*drsa_ref.step_sums(A.double(), C.double(), U.double(), K), M, K)[0])
    sums = torch.cat([Xq.reshape(-1), ssq]).float().cuda()
    Ud = U.cuda().contiguous()
    ws = torch.empty(int(L.check(lib.drsa_finish_workspace_bytes(d, d))), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    out = {}
    for flag in (0, 1):
        for update in (False, True):
            log = torch.zeros(4, device="cuda"); status = torch.zeros(4, dtype=torch.int32, device="cuda")
            Uo = torch.empty_like(Ud) if update else None
            L.check(lib.drsa_finish_step(_ptr(sums), M, _ptr(Ud), d, d, K, _ptr(Uo), None, None, _ptr(log), 0, 8, 1e-6, flag,
                                         _ptr(status), _ptr(ws), ws.numel(), s))
            out[(flag, update)] = float(log[0])
    want_corr = float(f_q) + float((grad_q * (U.double() - Uq)).sum())
    for update in (False, True):
        assert abs(out[(0, update)] - float(f_q)) / float(f_q) < 2e-6
        assert abs(out[(1, update)] - want_corr) / want_corr < 2e-6
    print(f"objective at fp16(U): rel err {abs(float(f_q) - f_true) / f_true:.2e}; first-order corrected: "
          f"{abs(out[(1, True)] - f_true) / f_true:.2e}")
    assert abs(out[(1, True)] - f_true) / f_true < 3e-6


def test_large_problem_properties(L):
    """cfg-2-sized rows (M = 640k, d = 256, K = 4): the tensor-core row pass agrees with the fp32 path
    on the same device, is additive over row shards, and a few steps increase the objective."""
    M, d, K = 640_000, 256, 4
    A, C = drsa_ref.synth_pairs(M, d, 20262, structured=False)
    U = drsa_ref.synth_U0(d, d, 4)
    Xt, st = _sums_gpu(L, A, C, U, K, "tc")
    Xf, sf = _sums_gpu(L, A, C, U.half().float(), K, "fp32")        # 'tc' evaluates at fp16(U)
    assert float(torch.linalg.norm(Xt - Xf) / torch.linalg.norm(Xf)) < 1e-4
    np.testing.assert_allclose(st.numpy(), sf.numpy(), rtol=1e-4)
    Xs, s_s = _sums_gpu(L, A, C, U, K, "tc_split")
    Xe, se = _sums_gpu(L, A, C, U, K, "fp32")
    assert float(torch.linalg.norm(Xs - Xe) / torch.linalg.norm(Xe)) < 1e-4
    np.testing.assert_allclose(s_s.numpy(), se.numpy(), rtol=1e-4)
    half = M // 2 + 37
    Xa, sa = _sums_gpu(L, A[:half], C[:half], U, K, "tc")
    Xb, sb = _sums_gpu(L, A[half:], C[half:], U, K, "tc")
    assert float(torch.linalg.norm(Xa + Xb - Xt) / torch.linalg.norm(Xt)) < 1e-5
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    opt = SubspaceOptimizer(U, A, C, None, num_concepts=K, precision="tc")
    opt.run(steps=8, save=False)
    assert np.all(np.diff(opt.obj_history) > 0)
    assert opt.last_status[1] == 0


def test_subspace_relevances_and_context_helpers():
    from cxai.xai.explain.explainer import compute_subspace_relevances
    from cxai.xai.drsa import preprocessing as pp
    g = torch.Generator().manual_seed(0)
    B, P, d, K = 5, 16, 64, 4
    a = torch.rand(B, P, d, generator=g); c = torch.randn(B, P, d, generator=g)
    U = drsa_ref.synth_U0(d)
    got = compute_subspace_relevances(a.cuda(), c.cuda(), U.cuda(), K).cpu()
    want = drsa_ref.subspace_relevances(a, c, U, K)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    # context vectors / normalisation / fused gather
    amap = torch.relu(torch.randn(6, 40, 4, 8, generator=g)); Rmap = torch.randn(6, 40, 4, 8, generator=g) * (amap > 0)
    cv = pp.compute_context_vectors(amap.cuda(), Rmap.cuda()).cpu()
    np.testing.assert_array_equal(cv.numpy(), drsa_ref.compute_context_vectors(amap, Rmap).numpy())
    nv = pp.normalize_vectors(amap.reshape(-1, 8).cuda()).cpu()
    np.testing.assert_allclose(nv.numpy(), drsa_ref.normalize_vectors(amap.reshape(-1, 8)).numpy(), rtol=2e-6)
    np.random.seed(3)
    idcs = pp.sample_spatial_locations(6, (4, 8), 5)
    np.random.seed(3)
    idcs2 = np.stack([np.random.choice(32, 5, replace=False) for _ in range(6)])
    np.testing.assert_array_equal(idcs, idcs2)
    for idx in (None, idcs):
        act, ctx = pp.gather_context_pairs(amap.cuda(), Rmap.cuda(), idx, normalize=True)
        av = drsa_ref.vectors_from_maps_all(amap) if idx is None else drsa_ref.vectors_from_maps_fixed(amap, idx)
        rv = drsa_ref.vectors_from_maps_all(Rmap) if idx is None else drsa_ref.vectors_from_maps_fixed(Rmap, idx)
        np.testing.assert_allclose(act.cpu().numpy(), drsa_ref.normalize_vectors(av).numpy(), rtol=3e-6, atol=1e-8)
        np.testing.assert_allclose(ctx.cpu().numpy(),
                                   drsa_ref.normalize_vectors(drsa_ref.compute_context_vectors(av, rv)).numpy(),
                                   rtol=3e-6, atol=1e-7)
    v_ref = pp.get_vectors_from_maps(amap, idcs, layout="reference")
    np.testing.assert_array_equal(v_ref.numpy(), drsa_ref.vectors_from_maps_ref(amap, idcs).numpy())


def test_preprocessing_kernels_match_reference_fixture(golden_dir):
    """compute_context_vectors / normalize_vectors (preprocessing.py:179-193, 219-231) on the GPU against the outputs of the
    reference's own functions (tests/golden/lrp_prep.npz, oracle/gen_golden_lrp.py `prep`)."""
    from cxai.xai.drsa import preprocessing as pp
    g = np.load(os.path.join(golden_dir, "lrp_prep.npz"))
    va, vr, c = (torch.from_numpy(g[k]).cuda() for k in ("va", "vr", "c"))
    np.testing.assert_array_equal(pp.compute_context_vectors(va, vr).cpu().numpy(), g["c"])         # one IEEE division
    gen = torch.Generator().manual_seed(int(g["seed"]))
    B, d, H, W = (int(g[k]) for k in ("B", "d", "H", "W"))
    amap = torch.relu(torch.randn(B, d, H, W, generator=gen))
    Rmap = torch.randn(B, d, H, W, generator=gen) * (amap > 0)
    np.testing.assert_array_equal(pp.compute_context_vectors(amap.cuda(), Rmap.cuda()).cpu().numpy(), g["c_maps"])
    np.testing.assert_allclose(pp.normalize_vectors(va).cpu().numpy(), g["na"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(pp.normalize_vectors(c).cpu().numpy(), g["nc"], rtol=2e-6, atol=1e-9)
    # the fused gather (corrected row layout) holds the same numbers as the reference's pipeline applied to the same rows
    act, ctx = pp.gather_context_pairs(amap.cuda(), Rmap.cuda(), g["idcs"], normalize=False)
    L = int(g["L"])
    rows_a = torch.stack([amap[b].flatten(1)[:, g["idcs"][b, l]] for b in range(B) for l in range(L)])
    rows_r = torch.stack([Rmap[b].flatten(1)[:, g["idcs"][b, l]] for b in range(B) for l in range(L)])
    np.testing.assert_array_equal(act.cpu().numpy(), rows_a.numpy())
    np.testing.assert_allclose(ctx.cpu().numpy(), (rows_r / (rows_a + 1e-7)).numpy(), rtol=1e-6, atol=0)


def test_batched_subset_objectives_match_obj_val_per_subset():
    """Prototype search (prototypes.py:98-119): the objective of every subset from one batched call equals the reference's
    obj_val evaluated subset by subset (oracle), also when the call is cut into several chunks."""
    from cxai.xai.drsa.prototypes import subset_objectives
    S, R, d, K = 37, 48, 64, 4
    A, C = drsa_ref.synth_pairs(S * R, d, 71)
    U = drsa_ref.synth_U0(d, d, 72)
    want = np.array([float(drsa_ref.obj_val(A[s * R:(s + 1) * R].double(), C[s * R:(s + 1) * R].double(), U.double(), K, d // K))
                     for s in range(S)])
    for max_rows in (1 << 21, 5 * R):
        got = subset_objectives(A.cuda(), C.cuda(), U.cuda(), S, R, K, max_rows=max_rows).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=2e-5)


@pytest.mark.parametrize("prec", ["auto", "tc_dc"])
def test_padded_production_shape_in_the_default_arithmetic(prec):
    """The reference's production split layer 19 (arch A: d = 100, K = 4 x 25, getdrsadata.py:72-73,119) at a row count where
    'auto' takes the tensor cores: the problem runs zero-padded to d = 128 with hi + lo row planes ('tc_hilo') or the deferred
    correction, and must follow the reference algorithm like any native shape."""
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    d, K, M, steps = 100, 4, 70000, 40
    A, C = drsa_ref.synth_pairs(M, d, 930)
    U0 = drsa_ref.synth_U0(d, seed=931)
    opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device="cuda", precision=prec)
    assert opt._pad is not None and opt.precision == ("tc_hilo" if prec == "auto" else "tc_dc")
    opt.run(steps=steps, save=False)
    objs_ref, U_ref = drsa_ref.run_autograd(A, C, U0, K, steps)
    rel = float(np.max(np.abs(opt.obj_history - objs_ref) / np.abs(objs_ref)))
    ang = drsa_ref.principal_angle(opt.U.cpu(), U_ref, K)
    assert rel < 1e-4 and ang < 1e-3, (rel, ang)
    assert tuple(opt.U.shape) == (d, d)


@pytest.mark.parametrize("d,m", [(64, 64), (128, 128), (256, 256), (256, 128), (512, 512)])
def test_qr_retraction_option_matches_householder_qr_with_sign_fix(d, m):
    """Non-default option (BASELINE north_star (3)): Q of the thin QR factorisation with diag(R) > 0."""
    from cxai.xai.drsa.drsa import orthogonalize
    g = torch.Generator().manual_seed(7 * d + m)
    Y = drsa_ref.synth_U0(d, m, 3) + 0.3 * torch.randn(d, m, generator=g) / d ** 0.5
    got = orthogonalize(Y.cuda(), method="qr").cpu().double()
    want = drsa_ref.orthogonalize_qr(Y.double())
    assert float((got - want).abs().max()) < 5e-6
    assert float((got.T @ got - torch.eye(m, dtype=torch.float64)).abs().max()) < 5e-6


def test_qr_retraction_gives_another_trajectory_than_the_reference():
    """retraction='qr' follows ITS oracle (same steps with a QR retraction) and, as SURVEY F1 states, not the reference."""
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    M, d, K, steps = 3000, 64, 4, 8
    A, C = drsa_ref.synth_pairs(M, d, 61)
    U0 = drsa_ref.synth_U0(d, d, 62)
    opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device="cuda", retraction="qr")
    opt.run(steps=steps, save=False)
    U, objs = U0.double(), []
    for _ in range(steps):
        o, g, _ = drsa_ref.step_closed_form(A.double(), C.double(), U, K)
        objs.append(float(o))
        U = drsa_ref.orthogonalize_qr(U + g)
    objs.append(float(drsa_ref.finish_from_sums(*drsa_ref.step_sums(A.double(), C.double(), U, K), M, K)[0]))
    np.testing.assert_allclose(opt.obj_history, np.asarray(objs), rtol=1e-4)
    assert float((opt.U.cpu().double() - U).abs().max()) < 1e-4
    objs_ref, U_ref = drsa_ref.run_autograd(A, C, U0, K, steps)
    assert np.max(np.abs(opt.obj_history - objs_ref) / objs_ref) > 1e-3          # a different trajectory
