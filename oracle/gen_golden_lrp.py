"""Stage-1 golden fixtures: the reference's OWN stage-1 code, unmodified, on top of the mini-zennit restatement.

TEST INFRASTRUCTURE ONLY.  Run in the build container as a script (``python oracle/gen_golden_lrp.py [case ...]``): it
puts /root/reference in front of ``sys.path`` so that ``cxai`` is the REFERENCE package (this repository's ``cxai`` is never
imported in this process), registers ``oracle/mini_zennit`` as ``zennit`` (the real package is unobtainable, see its
header) plus an empty ``librosa`` stub (imported by cxai/utils/dataloading.py:7, unused on this path), and calls

  * ``cxai.xai.drsa.preprocessing.get_intermediate``            (preprocessing.py:106-176)
  * ``cxai.xai.explain.attribute.compute_relevances``           (attribute.py:70-108)
  * ``cxai.xai.explain.explainer.HeatmapGenerator``             (explainer.py:15-177, with ``ProjectionModel``,
    ``SubspaceHook`` and ``get_class_composite`` under it)
  * ``cxai.xai.explain.explainer.compute_subspace_relevances``  (explainer.py:206-242)

on models built by the reference's ``VGGType`` constructor (create_model.py:8-171; only ``forward`` is overridden where
the hard-coded ``x.view(-1, 2048)`` of create_model.py:95 does not fit the configuration) with the reference's
rule maps (constants.py:27-51; getdrsadata.py:87-108 restated as a list because that script is not importable).

What is pinned by these fixtures: orchestration, hooks, seeds, model construction, projection layers -- everything that is
reference code.  What is not: the rule arithmetic inside mini-zennit (restated).  Every output is stored from an fp64 run of
the same code (``*_f64``, float32 storage) and from the fp32 run as shipped; ``noise_*`` is their norm-wise distance.
"""
from __future__ import annotations

import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("DRSA_REFERENCE_ROOT", "/root/reference")
sys.path = [REFERENCE_ROOT] + [p for p in sys.path if os.path.abspath(p or ".") != REPO] + [REPO]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import mini_zennit  # noqa: E402
from oracle import synth  # noqa: E402

mini_zennit.install()
sys.modules.setdefault("librosa", types.ModuleType("librosa"))

import cxai  # noqa: E402

assert os.path.abspath(cxai.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), cxai.__file__
from cxai.model.create_model import VGGType as RefVGGType  # noqa: E402
from cxai.xai.drsa.preprocessing import get_intermediate  # noqa: E402
from cxai.xai.drsa import preprocessing as ref_pp  # noqa: E402
from cxai.xai.explain.attribute import compute_relevances  # noqa: E402
from cxai.xai.explain.explainer import HeatmapGenerator, compute_subspace_relevances  # noqa: E402
from cxai.utils.constants import LRP_NAME_MAP_GTZAN, LRP_NAME_MAP_TOY  # noqa: E402
from zennit.rules import Epsilon, Gamma, WSquare  # noqa: E402
from zennit.composites import NameMapComposite  # noqa: E402
from zennit.canonizers import SequentialMergeBatchNorm  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


class FlatVGG(RefVGGType):
    """The reference's constructor, with the flatten size taken from the tensor (create_model.py:95 hard-codes 2048)."""

    def forward(self, x):
        x = self.features(x)
        return self.classifier(x.view(x.size(0), -1))


def name_map_6s(gamma=0.3, stab=1e-7):
    """getdrsadata.py:87-108."""
    return [(['features.0'], WSquare(stabilizer=stab)),
            (['features.3'], Gamma(gamma=gamma, stabilizer=stab)),
            (['features.7'], Gamma(gamma=gamma, stabilizer=stab)),
            (['features.10'], Gamma(gamma=gamma, stabilizer=stab)),
            (['features.14'], Gamma(gamma=gamma / 2, stabilizer=stab)),
            (['features.17'], Gamma(gamma=gamma / 2, stabilizer=stab)),
            (['features.21'], Gamma(gamma=gamma / 2, stabilizer=stab)),
            (['features.24'], Gamma(gamma=gamma / 2, stabilizer=stab)),
            (['features.28'], Gamma(gamma=gamma / 4, stabilizer=stab)),
            (['features.31'], Gamma(gamma=gamma / 4, stabilizer=stab)),
            (['classifier.0'], Epsilon(epsilon=stab)),
            (['classifier.4'], Epsilon(epsilon=stab)),
            (['classifier.8'], Epsilon(epsilon=stab))]


def _rel(a, b):
    """norm-wise relative distance per sample"""
    a, b = a.double().flatten(1), b.double().flatten(1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp(min=1e-300)).numpy()


def _both(fn, net, x):
    """Run ``fn(net, x)`` (dict of tensors) in fp32 as shipped and in fp64; returns the merged dict + noise figures."""
    out32 = fn(net, x.clone())
    out64 = fn(net.double(), x.double())
    net.float()
    out = {}
    for k in out32:
        if out32[k].numel() <= 4096:                       # the fp32 run as shipped is kept for small outputs only;
            out[k] = out32[k].detach().float().numpy()     # for the maps its distance to the fp64 run is recorded
        out[k + "_f64"] = out64[k].detach().float().numpy()
        out["noise_" + k] = _rel(out32[k].detach(), out64[k].detach())
    return out


def case_toy():
    net = synth.build_model(FlatVGG, "toy", seed=0, bn_seed=None)
    x = synth.synth_logmel(70, 64, 64, 20261)               # 2 minibatches (64 + 6), preprocessing.py:139
    comp = lambda: NameMapComposite(LRP_NAME_MAP_TOY)       # constants.py:40-51

    def run(net, x):
        o = {}
        for cls, onehot in ((0, False), (1, True)):
            a, R = get_intermediate(net, x, comp(), net.features[13], cls, one_hot_encoded=onehot)
            o[f"a_c{cls}"], o[f"R_c{cls}"] = a, R
        o["Rin_c0"] = compute_relevances(net, x[:8].clone(), comp(), class_idx=0)
        o["Rin_all"] = compute_relevances(net, x[:8].clone(), comp(), num_classes=2)       # attribute.py:148-158
        o["logits"] = net(x[:8]).detach()
        return o
    out = _both(run, net, x)
    out.update(model="toy", seed=0, x_seed=20261, N=70, wsum=synth.weight_checksum(net))
    return out


def case_archA_small():
    net = synth.build_model(FlatVGG, "archA_small", seed=0, bn_seed=1)
    x = synth.synth_logmel(6, 32, 64, 20262)
    comp = lambda: NameMapComposite(name_map_6s(), canonizers=[SequentialMergeBatchNorm()])

    def run(net, x):
        o = {"Rin_c3": compute_relevances(net, x.clone(), comp(), class_idx=3)}
        for layer in (19, 26, 33):                          # getdrsadata.py:119
            a, R = get_intermediate(net, x, comp(), net.features[layer], 3)
            o[f"a_l{layer}"], o[f"R_l{layer}"] = a, R
        o["logits"] = net(x).detach()
        return o
    out = _both(run, net, x)
    out.update(model="archA_small", seed=0, bn_seed=1, x_seed=20262, N=6, wsum=synth.weight_checksum(net))
    return out


def case_cfg2_full():
    """BASELINE cfg 2 CNN at full resolution (128 x 256, d = 256 at features[33]), 2 samples."""
    net = synth.build_model(FlatVGG, "cfg2", seed=0, bn_seed=1)
    x = synth.synth_logmel(2, 128, 256, 20263)
    comp = lambda: NameMapComposite(name_map_6s(), canonizers=[SequentialMergeBatchNorm()])

    def run(net, x):
        a, R = get_intermediate(net, x, comp(), net.features[33], 3)
        return {"a_l33": a, "R_l33": R, "Rin_c3": compute_relevances(net, x.clone(), comp(), class_idx=3),
                "logits": net(x).detach()}
    out = _both(run, net, x)
    out.update(model="cfg2", seed=0, bn_seed=1, x_seed=20263, N=2, wsum=synth.weight_checksum(net))
    return out


def case_archB():
    """The 3-second GTZAN model of pixelflipping/cpf.py:410-412 (arch B) at its own resolution with the reference's own rule
    map for it (constants.py:27-38), split at two of the layers cpf.py:141 uses."""
    net = synth.build_model(RefVGGType, "archB", seed=0, bn_seed=None)      # flat size 2048: the reference's forward fits
    x = synth.synth_logmel(3, 128, 128, 20266)
    comp = lambda: NameMapComposite(LRP_NAME_MAP_GTZAN)

    def run(net, x):
        o = {"Rin_c6": compute_relevances(net, x.clone(), comp(), class_idx=6)}
        for layer in (7, 13):
            a, R = get_intermediate(net, x, comp(), net.features[layer], 6)
            o[f"a_l{layer}"], o[f"R_l{layer}"] = a, R
        o["logits"] = net(x).detach()
        return o
    out = _both(run, net, x)
    out.update(model="archB", seed=0, x_seed=20266, N=3, wsum=synth.weight_checksum(net))
    return out


def _heatmaps(model_name, H, W, layer_idx, sample_class, name_map_fn, canon, N, x_seed, K=4):
    """HeatmapGenerator (explainer.py:15-177) for a signed-permutation U (projections exact) and a random orthogonal U."""
    net = synth.build_model(RefVGGType if model_name == "archA" else FlatVGG, model_name, seed=0,
                            bn_seed=1 if canon else None)
    x = synth.synth_logmel(N, H, W, x_seed)
    d = [m for m in list(net.features)[:layer_idx] if isinstance(m, torch.nn.Conv2d)][-1].out_channels
    out = dict(model=model_name, seed=0, x_seed=x_seed, N=N, K=K, layer_idx=layer_idx, d=d,
               sample_class=sample_class, wsum=synth.weight_checksum(net))
    if canon:
        out["bn_seed"] = 1
    for tag, U in (("perm", synth.signed_permutation(d, 5)), ("orth", synth.random_orthogonal(d, 6))):
        res = {}
        for dt in (torch.float32, torch.float64):
            net.to(dt)
            if canon:
                # get_class_composite builds its composite WITHOUT canonizers (explainer.py:203): the reference's users
                # merge the batch norms themselves; here the canonizer is applied around the call
                handles = SequentialMergeBatchNorm().apply(net)
            gen = HeatmapGenerator(net, U.to(dt), name_map_fn(), sample_class, num_concepts=K, layer_idx=layer_idx)
            gen.generate_subspace_heatmaps(x.to(dt).clone())
            res[dt] = {k: np.asarray(v) for k, v in gen.info.items() if k != "input"}
            if canon:
                for h in handles:
                    h.remove()
        net.float()
        for k, v in res[torch.float64].items():
            if k == "subspace_heatmaps" and tag == "orth" and v.size > 200000:
                # individual concept maps under a general U are only defined to ~1e-2 (see noise_orth_subspace_heatmaps:
                # where a ReLU output is exactly 0, a' = (aU)U^T is rounding noise and the Epsilon quotient lets part of
                # the relevance through with the sign of the noise); their sum is well defined and is what is stored
                out[f"{tag}_subspace_sum"] = v.sum(axis=1, keepdims=True).astype(np.float32)
                continue
            out[f"{tag}_{k}"] = v.astype(np.float32) if v.dtype.kind == "f" else v
        for k in ("standard_heatmaps", "subspace_heatmaps"):
            a, b = torch.from_numpy(res[torch.float32][k]), torch.from_numpy(res[torch.float64][k])
            out[f"noise_{tag}_{k}"] = _rel(a, b)
        # the per-instance concept relevances at the split layer through the reference's own function
        # (explainer.py:206-242) on the maps of this model: a, c = R/(a+1e-7)
        a, R = get_intermediate(net, x, NameMapComposite(name_map_fn(), canonizers=[SequentialMergeBatchNorm()] if canon
                                                         else None), net.features[layer_idx],
                                gen.class_idx)
        av = a.flatten(2).transpose(1, 2)
        cv = (R / (a + 1e-7)).flatten(2).transpose(1, 2)
        out[f"{tag}_Rk"] = compute_subspace_relevances(av, cv, U, K).numpy()
    return out


def case_heat_toy():
    return _heatmaps("toy16", 64, 64, 10, "class1", lambda: list(LRP_NAME_MAP_TOY), False, N=4, x_seed=20264)


def case_heat_archA():
    return _heatmaps("archA", 128, 256, 33, "rock", name_map_6s, True, N=2, x_seed=20265)   # N = 1 crashes explainer.py:175


def case_archB_early():
    """The other split layers of cpf.py:141 on arch B: features[1], [4] (d = 32) and [10] (d = 64).  The maps of the two early
    layers are large (32 x 128 x 128 per sample) and are stored at every 4th pixel in both directions."""
    net = synth.build_model(RefVGGType, "archB", seed=0, bn_seed=None)
    x = synth.synth_logmel(3, 128, 128, 20266)[:2]          # the first two samples of case_archB
    comp = lambda: NameMapComposite(LRP_NAME_MAP_GTZAN)

    def run(net, x):
        o = {}
        for layer, st in ((1, 4), (4, 4), (10, 1)):
            a, R = get_intermediate(net, x, comp(), net.features[layer], 6)
            o[f"a_l{layer}"], o[f"R_l{layer}"] = a[..., ::st, ::st].contiguous(), R[..., ::st, ::st].contiguous()
        return o
    out = _both(run, net, x)
    out.update(model="archB", seed=0, x_seed=20266, N=2, wsum=synth.weight_checksum(net))
    return out


def case_prep():
    """The small helpers between the LRP pass and the optimiser, the reference's own functions (preprocessing.py:179-256):
    sample_spatial_locations (global numpy RNG), get_vectors_from_maps (with its row scrambling),
    compute_context_vectors, normalize_vectors.  Not a `_both` case: outputs are stored as the fp32 run produced them."""
    g = torch.Generator().manual_seed(20269)
    B, d, H, W, L = 6, 40, 4, 8, 5
    amap = torch.relu(torch.randn(B, d, H, W, generator=g))
    Rmap = torch.randn(B, d, H, W, generator=g) * (amap > 0)
    np.random.seed(3)
    idcs = ref_pp.sample_spatial_locations(B, (H, W), L)                 # :196-216
    va, vr = ref_pp.get_vectors_from_maps(amap, idcs), ref_pp.get_vectors_from_maps(Rmap, idcs)      # :234-256
    c = ref_pp.compute_context_vectors(va, vr)                           # :179-193
    call = ref_pp.compute_context_vectors(amap, Rmap)
    return dict(seed=20269, np_seed=3, B=B, d=d, H=H, W=W, L=L, idcs=idcs, va=va.numpy(), vr=vr.numpy(), c=c.numpy(),
                c_maps=call.numpy(), na=ref_pp.normalize_vectors(va).numpy(), nc=ref_pp.normalize_vectors(c).numpy())


def case_hostutils():
    """Host-side utilities either side of the path, the reference's own functions: get_slice (utils/sound.py:8-44) on a seeded
    ~29.3 s clip at a reduced sample rate, get_best_run (utils/evaluation.py:107-141) on a result tree written here."""
    import tempfile
    from cxai.utils.sound import get_slice
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):       # plotting imports of evaluation.py:5-7, unused here
        sys.modules.setdefault(name, types.ModuleType(name))
    from cxai.utils.evaluation import get_best_run
    sr = 1000
    wav = torch.randn(1, int(29.3 * sr), generator=torch.Generator().manual_seed(20270))
    out = dict(sr=sr, wav_seed=20270, wav_len=wav.size(1))
    combos = [(3, 10), (6, 3), (6, 5), (3, 2)]
    out["slice_combos"] = np.asarray(combos)
    for sl, nc in combos:
        out[f"slices_{sl}_{nc}"] = get_slice(wav, sl, 7, nc, sr).numpy()
    out["slice_single"] = get_slice(wav, 6, 4, 1, sr).numpy()
    losses = {1: [0.10, 0.20, 0.31], 2: [0.12, 0.25, 0.42], 3: [0.11, 0.22, 0.40]}
    with tempfile.TemporaryDirectory() as tmp:
        for r, ls in losses.items():
            os.makedirs(os.path.join(tmp, f"run{r}"))
            with open(os.path.join(tmp, f"run{r}", "train_stats.csv"), "w") as f:       # the layout drsa.py:157-165 writes
                f.write(",loss\n" + "".join(f"{i},{v}\n" for i, v in enumerate(ls)))
        run, loss, crel, path, _ = get_best_run(tmp)
        out.update(best_run=run, best_loss=loss, best_dir=os.path.basename(path), n_concept_relevances=len(crel))
    out["tree_losses"] = np.asarray([losses[r] for r in (1, 2, 3)])
    return out


def case_heat_archB():
    """Concept heatmaps on arch B at the deepest split layer cpf.py:141 uses (features[13], d = 128)."""
    return _heatmaps("archB", 128, 128, 13, "rock", lambda: list(LRP_NAME_MAP_GTZAN), False, N=2, x_seed=20267)


CASES = {"toy": case_toy, "archA_small": case_archA_small, "archB": case_archB, "cfg2_full": case_cfg2_full, "heat_toy": case_heat_toy,
         "heat_archA": case_heat_archA, "heat_archB": case_heat_archB, "archB_early": case_archB_early, "prep": case_prep, "hostutils": case_hostutils}

if __name__ == "__main__":
    torch.set_num_threads(4)
    for name in (sys.argv[1:] or list(CASES)):
        out = CASES[name]()
        path = os.path.join(GOLD, f"lrp_{name}.npz")
        np.savez_compressed(path, **out)
        noise = {k: f"{float(np.max(v)):.1e}" for k, v in out.items() if k.startswith("noise_")}
        print(name, os.path.getsize(path), "bytes; fp32-vs-fp64 distance of the reference:", noise, flush=True)
