"""Golden result tree of the reference's multi-run driver ``drsa.main`` (/root/reference/cxai/xai/drsa/drsa.py:241-300),
UNMODIFIED, on seeded synthetic rows: numpy seed -> ``ortho_group.rvs(d)`` -> one cumulative column permutation per run ->
``SubspaceOptimizer.run`` -> ``run{r}/projection_matrix.pkl`` + ``run{r}/train_stats.csv``.

TEST INFRASTRUCTURE ONLY.  Run in the build container (``python -m oracle.gen_golden_main``); the output
``tests/golden/drsa_main.npz`` is committed (per run: the objective log and the saved U; the rows are regenerated from the
seed by the tests).
"""
from __future__ import annotations

import os
import pickle
import tempfile

import numpy as np
import torch

from oracle import drsa_ref
from oracle.ref_import import load_reference_drsa

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
M, D, K, STEPS, RUNS, SEED, ROW_SEED = 600, 32, 2, 12, 3, 1, 9


def generate():
    ref = load_reference_drsa()
    A, C = drsa_ref.synth_pairs(M, D, ROW_SEED)
    torch.set_num_threads(4)
    out = dict(M=M, d=D, K=K, steps=STEPS, runs=RUNS, seed=SEED, row_seed=ROW_SEED)
    with tempfile.TemporaryDirectory() as tmp:
        ref.main(A, C, tmp, num_concepts=K, steps=STEPS, runs=RUNS, seed=SEED, device="cpu")      # drsa.py:241-300
        for r in range(1, RUNS + 1):
            with open(os.path.join(tmp, f"run{r}", "projection_matrix.pkl"), "rb") as f:
                U = pickle.load(f)
            assert isinstance(U, np.ndarray) and U.dtype == np.float32 and U.shape == (D, D)
            lines = open(os.path.join(tmp, f"run{r}", "train_stats.csv")).read().splitlines()
            out[f"header"] = lines[0]
            out[f"loss_run{r}"] = np.asarray([float(l.split(",")[1]) for l in lines[1:]])
            out[f"U_run{r}"] = U
    return out


if __name__ == "__main__":
    o = generate()
    path = os.path.join(GOLD, "drsa_main.npz")
    np.savez_compressed(path, **o)
    print(path, os.path.getsize(path), "bytes;", {k: (v.shape if hasattr(v, "shape") else v) for k, v in o.items()})
