"""Generate tests/golden/drsa_*.npz by running the UNMODIFIED reference drsa.py.

Run in the build container (``python -m oracle.gen_golden``); the GPU box has no
/root/reference, so the outputs are committed as small fixtures.  Every array in a
fixture is an output of the reference's own code on the stored / seeded inputs.
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


def _case(ref, name, M, d, m, K, steps, seed, store_inputs, light=False):
    torch.manual_seed(0)
    A, C = drsa_ref.synth_pairs(M, d, seed)
    U0 = drsa_ref.synth_U0(d, m, seed + 1)
    d_k = m // K
    out = {"M": M, "d": d, "m": m, "K": K, "steps": steps, "seed": seed}
    if store_inputs:
        out.update(A=A.numpy(), C=C.numpy())
    if not light:       # light fixtures regenerate U0 from the seed (drsa_ref.synth_U0(d, m, seed + 1)) as well
        out["U0"] = U0.numpy()
    out["in_checksum"] = np.array([A.double().sum().item(), C.double().sum().item(),
                                   (A.double() * C.double()).sum().item()])
    # objective + autograd gradient at U0 (drsa.py:91-100)
    U = U0.clone().requires_grad_(True)
    obj = ref.SubspaceOptimizer.obj_val(A, C, U, ref.objective_fn, K, d_k)
    obj.backward()
    out["obj0"] = obj.detach().numpy()
    if not light:
        out["grad0"] = U.grad.numpy()
        # retraction (drsa.py:201-221)
        out["U1"] = ref.orthogonalize(U0 + U.grad).numpy()
    if m == d:
        # full loop through the reference class (drsa.py:76-120) incl. its file outputs
        with tempfile.TemporaryDirectory() as tmp:
            opt = ref.SubspaceOptimizer(U0.clone(), A, C, tmp, num_concepts=K, device="cpu")
            opt.run(steps=steps)
            with open(os.path.join(tmp, "projection_matrix.pkl"), "rb") as f:
                out["U_final"] = pickle.load(f)
            csv = open(os.path.join(tmp, "train_stats.csv")).read().splitlines()
            out["csv_header"] = np.array(csv[0])
            out["csv_rows"] = len(csv) - 1
            out["objs"] = np.array([float(l.split(",")[1]) for l in csv[1:]])
    else:
        # rectangular U: only the static functions apply
        Ucur, objs = U0.clone(), []
        for _ in range(steps):
            Ucur = Ucur.detach().requires_grad_(True)
            o = ref.SubspaceOptimizer.obj_val(A, C, Ucur, ref.objective_fn, K, d_k)
            o.backward()
            objs.append(float(o))
            Ucur = ref.orthogonalize(Ucur.detach() + Ucur.grad)
        objs.append(float(ref.SubspaceOptimizer.obj_val(A, C, Ucur.detach(), ref.objective_fn, K, d_k)))
        out["U_final"] = Ucur.detach().numpy()
        out["objs"] = np.array(objs)
    np.savez_compressed(os.path.join(GOLD, f"drsa_{name}.npz"), **out)
    print(name, "obj0", float(out["obj0"]), "obj_end", out["objs"][-1], "bytes",
          os.path.getsize(os.path.join(GOLD, f"drsa_{name}.npz")))


def main(only=None):
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference_drsa()
    torch.set_num_threads(1)   # deterministic summation order for the fixtures
    if only == "d512":         # cfg-4 width (added later; the other fixtures are left untouched)
        _case(ref, "d512", M=8192, d=512, m=512, K=8, steps=8, seed=17, store_inputs=False, light=True)
        return
    _case(ref, "tiny", M=384, d=32, m=32, K=4, steps=20, seed=11, store_inputs=True)
    _case(ref, "ragged", M=333, d=32, m=32, K=2, steps=10, seed=12, store_inputs=True)
    _case(ref, "toy64", M=4000, d=64, m=64, K=4, steps=60, seed=13, store_inputs=False)
    _case(ref, "d128", M=6000, d=128, m=128, K=4, steps=30, seed=14, store_inputs=False)
    _case(ref, "d256", M=8192, d=256, m=256, K=4, steps=12, seed=15, store_inputs=False)
    _case(ref, "rect", M=3000, d=64, m=32, K=2, steps=20, seed=16, store_inputs=False)
    # scalar helpers (drsa.py:171-182, 224-238)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(50, 4, generator=g)
    np.savez_compressed(os.path.join(GOLD, "drsa_fmean.npz"), x=x.numpy(),
                        f2=ref.generalized_fmean(x, 2).numpy(),
                        f05=ref.generalized_fmean(x, 0.5).numpy(),
                        obj=ref.objective_fn(x).numpy())


if __name__ == "__main__":
    import sys
    main(sys.argv[1] if len(sys.argv) > 1 else None)
