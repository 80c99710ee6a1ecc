"""Stage 1 on the GPU against fixtures produced by the reference's OWN code (oracle/gen_golden_lrp.py): the reference's
``get_intermediate`` (preprocessing.py:106-176), ``compute_relevances`` (attribute.py:70-108), ``HeatmapGenerator``
(explainer.py:15-177) and ``compute_subspace_relevances`` (explainer.py:206-242), unmodified, on models built by the
reference's ``VGGType`` constructor, running on the mini-zennit restatement (the rule arithmetic under them is the only
restated layer).  Metric: norm-wise relative error per sample <= 1e-4 against the fp64 run of the reference code; where the
reference's own fp32 run is further than that from its fp64 run (``noise_*`` in the fixture: max-pool near-ties), the bound
is 3x that distance."""
import os

import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _rel(got, want):
    got, want = torch.as_tensor(got).double().cpu().flatten(1), torch.as_tensor(want).double().flatten(1)
    return ((got - want).norm(dim=1) / want.norm(dim=1).clamp(min=1e-300)).numpy()


def _check(got, g, key, label=""):
    noise = np.asarray(g["noise_" + key], dtype=np.float64).reshape(-1) if "noise_" + key in g.files else np.zeros(1)
    err = _rel(got, g[key + "_f64"] if key + "_f64" in g.files else g[key])
    bound = np.maximum(TOL, 3.0 * noise)
    print(f"{label}{key}: max err {err.max():.2e} (reference fp32-vs-fp64 noise {noise.max():.1e})")
    assert np.all(err < bound), (key, err, bound)


def _load(golden_dir, name):
    from cxai.model.create_model import VGGType
    g = np.load(os.path.join(golden_dir, f"lrp_{name}.npz"))
    net = synth.build_model(VGGType, str(g["model"]), int(g["seed"]), int(g["bn_seed"]) if "bn_seed" in g.files else None)
    np.testing.assert_allclose(synth.weight_checksum(net), g["wsum"], rtol=1e-13)    # same model as the reference process built
    return g, net


def _composite(bn: bool):
    from cxai.utils.constants import LRP_NAME_MAP_TOY, lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    return NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()]) if bn else \
        NameMapComposite(LRP_NAME_MAP_TOY)


def test_toy_cfg1_against_reference_fixture(golden_dir):
    from cxai.xai.drsa.preprocessing import get_intermediate
    from cxai.xai.explain.attribute import compute_relevances
    g, net = _load(golden_dir, "toy")
    x = synth.synth_logmel(int(g["N"]), 64, 64, int(g["x_seed"]))
    comp = _composite(False)
    for cls, onehot in ((0, False), (1, True)):
        a, R = get_intermediate(net, x, comp, net.features[13], cls, one_hot_encoded=onehot)
        _check(a, g, f"a_c{cls}")
        _check(R, g, f"R_c{cls}")
    _check(compute_relevances(net, x[:8], comp, class_idx=0), g, "Rin_c0")
    _check(compute_relevances(net, x[:8], comp, num_classes=2), g, "Rin_all")


def test_bn_model_against_reference_fixture(golden_dir):
    from cxai.xai.drsa.preprocessing import get_intermediate
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.xai.explain import lrp_engine
    g, net = _load(golden_dir, "archA_small")
    x = synth.synth_logmel(int(g["N"]), 32, 64, int(g["x_seed"])).cuda()
    comp = _composite(True)
    assert lrp_engine._plan(net, comp, x.device)._tc_stack_ok(x)            # the tcgen05 stack is what is being compared
    _check(compute_relevances(net, x, comp, class_idx=3), g, "Rin_c3")
    for layer in (19, 26, 33):
        a, R = get_intermediate(net, x, comp, net.features[layer], 3)
        _check(a, g, f"a_l{layer}")
        _check(R, g, f"R_l{layer}")
    assert not lrp_engine._plan(net, comp, x.device).tc_failed()


def test_arch_b_against_reference_fixture(golden_dir):
    """The reference's 3-second GTZAN model (cpf.py:410-412: 32/32/64/64/128 filters, one conv per block, no BatchNorm,
    128 x 128 input, square pools) under LRP_NAME_MAP_GTZAN (constants.py:27-38), split at features[7] and features[13]."""
    from cxai.utils.constants import LRP_NAME_MAP_GTZAN
    from cxai.xai.drsa.preprocessing import get_intermediate
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.xai.explain.rules import NameMapComposite
    g, net = _load(golden_dir, "archB")
    x = synth.synth_logmel(int(g["N"]), 128, 128, int(g["x_seed"])).cuda()
    comp = NameMapComposite(LRP_NAME_MAP_GTZAN)
    _check(compute_relevances(net, x, comp, class_idx=6), g, "Rin_c6")
    for layer, d, hw in ((7, 64, 32), (13, 128, 8)):
        a, R = get_intermediate(net, x, comp, net.features[layer], 6)
        assert tuple(a.shape) == (3, d, hw, hw)
        _check(a, g, f"a_l{layer}")
        _check(R, g, f"R_l{layer}")


def test_arch_b_early_split_layers_against_reference_fixture(golden_dir):
    """features[1], [4] (d = 32) and [10] (d = 64): the remaining split layers of cpf.py:141 (early maps compared at stride 4,
    as stored)."""
    from cxai.utils.constants import LRP_NAME_MAP_GTZAN
    from cxai.xai.drsa.preprocessing import get_intermediate
    from cxai.xai.explain.rules import NameMapComposite
    g, net = _load(golden_dir, "archB_early")
    x = synth.synth_logmel(3, 128, 128, int(g["x_seed"]))[:int(g["N"])].cuda()
    comp = NameMapComposite(LRP_NAME_MAP_GTZAN)
    for layer, st, d, hw in ((1, 4, 32, 128), (4, 4, 32, 64), (10, 1, 64, 16)):
        a, R = get_intermediate(net, x, comp, net.features[layer], 6)
        assert tuple(a.shape) == (2, d, hw, hw)
        _check(a[..., ::st, ::st], g, f"a_l{layer}")
        _check(R[..., ::st, ::st], g, f"R_l{layer}")


def test_cfg2_cnn_full_resolution_against_reference_fixture(golden_dir):
    """BASELINE cfg 2 CNN (128 x 256 log-mel, d = 256 at features[33]) on the tensor-core stack vs the reference's code."""
    from cxai.xai.drsa.preprocessing import get_intermediate
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.xai.explain import lrp_engine
    g, net = _load(golden_dir, "cfg2_full")
    x = synth.synth_logmel(int(g["N"]), 128, 256, int(g["x_seed"])).cuda()
    comp = _composite(True)
    plan = lrp_engine._plan(net, comp, x.device)
    assert plan._tc_stack_ok(x)
    a, R = get_intermediate(net, x, comp, net.features[33], 3)
    assert a.shape == (2, 256, 8, 8)
    _check(a, g, "a_l33")
    _check(R, g, "R_l33")
    _check(compute_relevances(net, x, comp, class_idx=3), g, "Rin_c3")
    assert plan.use_tc and not plan.tc_failed()
    # context vectors of the hot path: c = R / (a + 1e-7) (preprocessing.py:193) on the reference's maps
    from cxai.xai.drsa.preprocessing import compute_context_vectors
    c = compute_context_vectors(a, R)
    af, Rf = torch.from_numpy(g["a_l33_f64"]).double(), torch.from_numpy(g["R_l33_f64"]).double()
    assert _rel(c, Rf / (af + 1e-7)).max() < 5e-4


@pytest.mark.parametrize("case", ["heat_toy", "heat_archA", "heat_archB"])
def test_heatmaps_against_reference_fixture(golden_dir, case):
    """HeatmapGenerator on the CUDA engine vs the reference's HeatmapGenerator (K + 1 clones through its ProjectionModel)."""
    from cxai.utils.constants import LRP_NAME_MAP_GTZAN, LRP_NAME_MAP_TOY, lrp_name_map_6s
    from cxai.xai.explain.explainer import HeatmapGenerator, compute_subspace_relevances
    from cxai.xai.explain.rules import SequentialMergeBatchNorm
    from cxai.xai.drsa.preprocessing import get_intermediate
    g, net = _load(golden_dir, case)
    K, layer_idx, d, N = int(g["K"]), int(g["layer_idx"]), int(g["d"]), int(g["N"])
    bn = case == "heat_archA"
    H, W = {"heat_toy": (64, 64), "heat_archA": (128, 256), "heat_archB": (128, 128)}[case]
    x = synth.synth_logmel(N, H, W, int(g["x_seed"]))
    nm = {"heat_toy": LRP_NAME_MAP_TOY, "heat_archA": lrp_name_map_6s(), "heat_archB": LRP_NAME_MAP_GTZAN}[case]
    for tag, U in (("perm", synth.signed_permutation(d, 5)), ("orth", synth.random_orthogonal(d, 6))):
        gen = HeatmapGenerator(net, U, nm, str(g["sample_class"]), num_concepts=K, layer_idx=layer_idx, device="cuda",
                               canonizers=[SequentialMergeBatchNorm()] if bn else ())
        gen.generate_subspace_heatmaps(x)
        _check(gen.info["standard_heatmaps"], g, f"{tag}_standard_heatmaps", label=case + " ")
        sub = gen.info["subspace_heatmaps"]
        if f"{tag}_subspace_heatmaps" in g.files:
            want = g[f"{tag}_subspace_heatmaps"]
            if tag == "perm":                                            # projections exact: every concept map is pinned
                np.testing.assert_array_equal(gen.info["mask"], g[f"{tag}_mask"])
                for j in range(K):
                    assert _rel(sub[:, j], want[:, j]).max() < TOL
                np.testing.assert_allclose(gen.info["subspace_relevances"], g[f"{tag}_subspace_relevances"], rtol=2e-3,
                                           atol=1e-4 * float(np.abs(g[f"{tag}_subspace_relevances"]).max()))
            ssum = want.sum(axis=1, keepdims=True)
        else:
            ssum = g[f"{tag}_subspace_sum"]
        # a general U leaves single concept maps defined to ~1e-2 only (noise_orth_subspace_heatmaps in the fixture: the
        # reference's own fp32 and fp64 runs differ by that much); their sum is pinned
        err = _rel(sub.sum(axis=1, keepdims=True), ssum)
        assert err.max() < max(TOL, 3 * float(np.max(g[f"noise_{tag}_standard_heatmaps"]))), err
        # per-instance concept relevances at the split layer (explainer.py:206-242)
        cls = gen.class_idx
        from cxai.xai.explain.rules import NameMapComposite
        comp = NameMapComposite(nm, canonizers=[SequentialMergeBatchNorm()] if bn else [])
        a, R = get_intermediate(net, x, comp, net.features[layer_idx], cls)
        av, cv = a.flatten(2).transpose(1, 2).contiguous(), (R / (a + 1e-7)).flatten(2).transpose(1, 2).contiguous()
        Rk = compute_subspace_relevances(av, cv, U.cuda(), K).cpu().numpy()
        np.testing.assert_allclose(Rk, g[f"{tag}_Rk"], rtol=2e-3, atol=2e-4 * float(np.abs(g[f"{tag}_Rk"]).max()))
