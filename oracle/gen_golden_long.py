"""Long-horizon golden trajectories: the UNMODIFIED reference ``SubspaceOptimizer.run`` at its own default horizon
(``steps=2000``, /root/reference/cxai/xai/drsa/drsa.py:76) on seeded synthetic rows.

TEST INFRASTRUCTURE ONLY.  Run in the build container (``python -m oracle.gen_golden_long [case ...]``); the outputs
``tests/golden/drsa_long_<case>.npz`` are committed (inputs are regenerated from the seed, only the reference's outputs
are stored: the objective at every step, U after every ``chunk`` steps and at the end).

The reference class is driven exactly as a user would: ``opt.run(chunk)`` is called ``steps/chunk`` times on the same
object (``run`` continues from ``self.U``, drsa.py:102), which yields U snapshots without touching the reference code;
the last objective of a chunk is the first of the next, so the stitched log equals that of one ``run(steps)``.

Each fixture also records the reference's distance to ITSELF at the horizon (``self_angle``, ``self_rel``): the same
reference code run with a different intra-op thread count (another summation order inside its GEMMs).  That is the noise
floor any other fp32 implementation of the same mathematics can be expected to reach.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

from oracle import drsa_ref
from oracle.ref_import import load_reference_drsa

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

#         name            M       d    K  steps chunk seed structured  alt-threads (0 = no self-distance run)
CASES = {
    "cfg1":        (16000,   64,  4, 2000, 500, 101, True, 1),      # BASELINE cfg 1: N = 1k x P = 16, d = 64, K = 4
    "d128_m8k":    (8192,   128,  4, 2000, 500, 102, True, 1),
    "d128_m64k":   (65536,  128,  4, 2000, 500, 103, True, 1),
    "d256_m8k":    (8192,   256,  4, 2000, 500, 104, True, 0),
    "d256_m64k":   (65536,  256,  4, 2000, 500, 105, True, 0),
    "d128_m256k":  (262144, 128,  4, 2000, 500, 106, True, 0),
    "d256_m256k":  (262144, 256,  4, 2000, 500, 107, True, 0),
    "d128_m8k_unstructured": (8192, 128, 4, 2000, 500, 108, False, 0),
}


def _run_reference(ref, A, C, U0, K, steps, chunk, threads):
    torch.set_num_threads(threads)
    with tempfile.TemporaryDirectory() as tmp:
        opt = ref.SubspaceOptimizer(U0.clone(), A, C, tmp, num_concepts=K, device="cpu")
        objs, snaps = [], []
        for c in range(steps // chunk):
            opt.run(steps=chunk)                                  # drsa.py:76-120, unmodified
            csv = open(os.path.join(tmp, "train_stats.csv")).read().splitlines()
            o = [float(l.split(",")[1]) for l in csv[1:]]
            assert len(o) == chunk + 1
            if objs:
                assert abs(objs[-1] - o[0]) <= 1e-6 * abs(o[0])   # final evaluation == first of the next chunk
                objs.extend(o[1:])
            else:
                objs.extend(o)
            snaps.append(opt.U.detach().clone().numpy())
    return np.asarray(objs), snaps


def make(name):
    M, d, K, steps, chunk, seed, structured, alt = CASES[name]
    ref = load_reference_drsa()
    A, C = drsa_ref.synth_pairs(M, d, seed, structured=structured)
    U0 = drsa_ref.synth_U0(d, d, seed + 1)
    t0 = time.time()
    objs, snaps = _run_reference(ref, A, C, U0, K, steps, chunk, threads=8)
    out = dict(M=M, d=d, m=d, K=K, steps=steps, chunk=chunk, seed=seed, structured=int(structured), threads=8,
               objs=objs, U_snaps=np.stack(snaps[:-1]).astype(np.float32), U_final=snaps[-1].astype(np.float32),
               in_checksum=np.array([A.double().sum().item(), C.double().sum().item(),
                                     (A.double() * C.double()).sum().item()]))
    if alt:
        objs2, snaps2 = _run_reference(ref, A, C, U0, K, steps, chunk, threads=alt)
        out["self_rel"] = float(np.max(np.abs(objs2 - objs) / np.abs(objs)))
        out["self_angle"] = float(drsa_ref.principal_angle(snaps2[-1], snaps[-1], K))
        out["self_angle_snaps"] = np.array([drsa_ref.principal_angle(a, b, K) for a, b in zip(snaps2, snaps)])
    path = os.path.join(GOLD, f"drsa_long_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: obj {objs[0]:.6f} -> {objs[-1]:.6f}; self rel {out.get('self_rel')}, self angle "
          f"{out.get('self_angle')} ({out.get('self_angle_snaps')}); {time.time() - t0:.0f} s, "
          f"{os.path.getsize(path)} bytes", flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        make(n)
