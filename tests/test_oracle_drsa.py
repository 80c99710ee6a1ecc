"""Pins the CPU oracle (oracle/drsa_ref.py) against golden vectors produced by the
UNMODIFIED reference drsa.py (oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import drsa_ref

CASES = ["tiny", "ragged", "toy64", "d128", "d256", "rect"]


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"drsa_{name}.npz"))
    M, d, m, K = int(g["M"]), int(g["d"]), int(g["m"]), int(g["K"])
    if "A" in g.files:
        A, C = torch.from_numpy(g["A"]), torch.from_numpy(g["C"])
    else:
        A, C = drsa_ref.synth_pairs(M, d, int(g["seed"]))
    chk = np.array([A.double().sum().item(), C.double().sum().item(), (A.double() * C.double()).sum().item()])
    np.testing.assert_allclose(chk, g["in_checksum"], rtol=1e-12)   # seeded inputs reproduce
    U0 = torch.from_numpy(g["U0"]) if "U0" in g.files else drsa_ref.synth_U0(d, m, int(g["seed"]) + 1)   # light fixtures
    return g, A, C, U0, K


@pytest.mark.parametrize("name", CASES)
def test_objective_and_gradient_match_reference(golden_dir, name):
    g, A, C, U0, K = _load(golden_dir, name)
    torch.set_num_threads(1)
    obj, grad, U1 = drsa_ref.step_autograd(A, C, U0, K)
    np.testing.assert_allclose(obj.numpy(), g["obj0"], rtol=1e-6)
    np.testing.assert_allclose(grad.numpy(), g["grad0"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(U1.numpy(), g["U1"], atol=2e-6)
    # closed form (what the CUDA kernels implement) in fp64 agrees with the reference's autograd
    o64, g64, _ = drsa_ref.step_closed_form(A.double(), C.double(), U0.double(), K)
    assert abs(float(o64) - float(g["obj0"])) / float(g["obj0"]) < 2e-6
    rel = np.linalg.norm(g64.numpy() - g["grad0"]) / np.linalg.norm(g["grad0"])
    assert rel < 5e-6, rel


@pytest.mark.parametrize("name", CASES + ["d512"])
def test_trajectory_matches_reference(golden_dir, name):
    g, A, C, U0, K = _load(golden_dir, name)
    torch.set_num_threads(1)
    steps = int(g["steps"])
    objs, U = drsa_ref.run_autograd(A, C, U0, K, steps)
    assert len(objs) == steps + 1 == len(g["objs"])              # drsa.py:104,117
    np.testing.assert_allclose(objs, g["objs"], rtol=2e-5)
    assert drsa_ref.principal_angle(U, g["U_final"], K) < 1e-4
    # fp64 closed form stays within the north-star tolerance of the fp32 reference
    if name == "d512":          # keep the CPU suite short: the fp64 twin of the widest case is covered by d256
        return
    objs64, U64 = drsa_ref.run_closed_form(A, C, U0, K, steps)
    assert np.max(np.abs(objs64 - g["objs"]) / g["objs"]) < 1e-4
    assert drsa_ref.principal_angle(U64, g["U_final"], K) < 1e-3
    # retraction invariant
    UtU = U.T @ U
    assert float((UtU - torch.eye(UtU.shape[0])).abs().max()) < 5e-6


def test_reference_file_outputs_pinned(golden_dir):
    g = np.load(os.path.join(golden_dir, "drsa_tiny.npz"))
    assert str(g["csv_header"]) == ",loss"                        # drsa.py:157-163 (pandas index col)
    assert int(g["csv_rows"]) == int(g["steps"]) + 1
    assert g["U_final"].dtype == np.float32                       # drsa.py:165-168


def test_fmean(golden_dir):
    g = np.load(os.path.join(golden_dir, "drsa_fmean.npz"))
    x = torch.from_numpy(g["x"])
    np.testing.assert_allclose(drsa_ref.generalized_fmean(x, 2).numpy(), g["f2"], rtol=1e-6)
    np.testing.assert_allclose(drsa_ref.generalized_fmean(x, 0.5).numpy(), g["f05"], rtol=1e-6)
    np.testing.assert_allclose(drsa_ref.objective_fn(x).numpy(), g["obj"], rtol=1e-6)


def test_sums_are_additive_over_row_shards():
    A, C = drsa_ref.synth_pairs(1000, 32, 7)
    U = drsa_ref.synth_U0(32)
    X, ss = drsa_ref.step_sums(A.double(), C.double(), U.double(), 4)
    X2 = torch.zeros_like(X); ss2 = torch.zeros_like(ss)
    for lo, hi in [(0, 300), (300, 301), (301, 1000)]:
        x, s = drsa_ref.step_sums(A[lo:hi].double(), C[lo:hi].double(), U.double(), 4)
        X2 += x; ss2 += s
    np.testing.assert_allclose(X2.numpy(), X.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ss2.numpy(), ss.numpy(), rtol=1e-12)


def test_vectors_from_maps_modes():
    maps = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 3, 2, 2)
    idcs = np.array([[0, 3], [1, 2]])
    fixed = drsa_ref.vectors_from_maps_fixed(maps, idcs)
    assert fixed.shape == (4, 3)
    np.testing.assert_array_equal(fixed[0].numpy(), maps[0, :, 0, 0].numpy())
    np.testing.assert_array_equal(fixed[3].numpy(), maps[1, :, 1, 0].numpy())
    scr = drsa_ref.vectors_from_maps_ref(maps, idcs)       # reference's scrambled layout
    assert scr.shape == (4, 3)
    assert not np.array_equal(scr.numpy(), fixed.numpy())
    allp = drsa_ref.vectors_from_maps_all(maps)
    np.testing.assert_array_equal(allp[5].numpy(), maps[1, :, 0, 1].numpy())


def test_port_matches_reference_at_the_full_horizon(golden_dir):
    """The oracle port over the reference's own 2 000 steps (drsa.py:76) on BASELINE cfg 1 (M = 16 000, d = 64, K = 4)
    against the golden trajectory of the unmodified reference (oracle/gen_golden_long.py).  The reference's distance to
    itself under another thread count is stored in the fixture; the port must be as close."""
    g = np.load(os.path.join(golden_dir, "drsa_long_cfg1.npz"))
    M, d, K, steps = int(g["M"]), int(g["d"]), int(g["K"]), int(g["steps"])
    A, C = drsa_ref.synth_pairs(M, d, int(g["seed"]))
    U0 = drsa_ref.synth_U0(d, d, int(g["seed"]) + 1)
    torch.set_num_threads(4)
    objs, U = drsa_ref.run_autograd(A, C, U0, K, steps)
    assert np.max(np.abs(objs - g["objs"]) / g["objs"]) < 1e-4
    ang = drsa_ref.principal_angle(U, g["U_final"], K)
    assert ang < max(1e-4, 5 * float(g["self_angle"])), ang
    assert float(g["self_angle"]) < 1e-3 and float(g["self_rel"]) < 1e-4


def test_multi_run_driver_sequence_matches_reference_result_tree(golden_dir):
    """The start matrices of the reference's main (drsa.py:265-283: numpy seed, ortho_group.rvs(d), one compounding column
    permutation per run) restated here, pushed through the oracle's trajectory, against the files the reference's own main
    wrote (oracle/gen_golden_main.py)."""
    from scipy.stats import ortho_group
    g = np.load(os.path.join(golden_dir, "drsa_main.npz"))
    M, d, K, steps, runs = (int(g[k]) for k in ("M", "d", "K", "steps", "runs"))
    A, C = drsa_ref.synth_pairs(M, d, int(g["row_seed"]))
    np.random.seed(int(g["seed"]))
    U = ortho_group.rvs(d)
    for r in range(1, runs + 1):
        U = U[:, np.random.permutation(d)]
        objs, Uf = drsa_ref.run_autograd(A, C, torch.tensor(U, dtype=torch.float32), K, steps)
        np.testing.assert_allclose(objs, g[f"loss_run{r}"], rtol=2e-6)
        assert drsa_ref.principal_angle(Uf, torch.from_numpy(g[f"U_run{r}"]), K) < 1e-5
