"""Pins the stage-1 CPU oracle (oracle/lrp_ref.py) and the mini-zennit restatement against fixtures produced by the
reference's OWN stage-1 code (oracle/gen_golden_lrp.py: get_intermediate, compute_relevances, HeatmapGenerator,
compute_subspace_relevances of /root/reference run unmodified on top of oracle/mini_zennit).  CPU only.

What these fixtures pin: orchestration, seeds, hooks, model construction, projection layers (reference code).
What they cannot pin: zennit's rule arithmetic (restated, package unobtainable) -- see oracle/mini_zennit/__init__.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import lrp_ref, synth, drsa_ref
from cxai.model.create_model import VGGType
from cxai.utils.constants import LRP_NAME_MAP_GTZAN, LRP_NAME_MAP_TOY, lrp_name_map_6s

TOL = 1e-4


def _rel(got, want):
    got, want = torch.as_tensor(got).double().flatten(1), torch.as_tensor(want).double().flatten(1)
    return ((got - want).norm(dim=1) / want.norm(dim=1).clamp(min=1e-300)).numpy()


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"lrp_{name}.npz"))
    net = synth.build_model(VGGType, str(g["model"]), int(g["seed"]), int(g["bn_seed"]) if "bn_seed" in g.files else None)
    # the product's model class builds bit-identical parameters from the same seed as the reference's constructor
    np.testing.assert_allclose(synth.weight_checksum(net), g["wsum"], rtol=1e-13)
    return g, net


def test_toy_maps_and_relevances_match_reference_fixture(golden_dir):
    g, net = _load(golden_dir, "toy")
    x = synth.synth_logmel(int(g["N"]), 64, 64, int(g["x_seed"]))
    for cls, onehot in ((0, False), (1, True)):
        a, R = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_TOY, net.features[13], cls, one_hot_encoded=onehot)
        assert _rel(a, g[f"a_c{cls}_f64"]).max() < 1e-6
        assert _rel(R, g[f"R_c{cls}_f64"]).max() < 1e-6
    o = lrp_ref.lrp_pass(net, x[:8], LRP_NAME_MAP_TOY, lrp_ref.output_modifier(0))
    assert _rel(o["R_input"], g["Rin_c0_f64"]).max() < 1e-6
    o = lrp_ref.lrp_pass(net, x[:8], LRP_NAME_MAP_TOY, lrp_ref.output_modifier(None, 2))
    assert _rel(o["R_input"], g["Rin_all_f64"]).max() < 1e-6
    np.testing.assert_allclose(o["logits"].numpy(), g["logits_f64"], rtol=1e-5, atol=1e-7)
    # the reference's own fp32 run stays within the tolerance of its fp64 run on this model
    assert max(float(g[k].max()) for k in g.files if k.startswith("noise_")) < TOL


def test_bn_model_matches_reference_fixture(golden_dir):
    g, net = _load(golden_dir, "archA_small")
    x = synth.synth_logmel(int(g["N"]), 32, 64, int(g["x_seed"]))
    nm = lrp_name_map_6s()
    o = lrp_ref.lrp_pass(net, x, nm, lrp_ref.output_modifier(3))
    assert _rel(o["R_input"], g["Rin_c3_f64"]).max() < 1e-6
    for layer in (19, 26, 33):
        a, R = lrp_ref.get_intermediate(net, x, nm, net.features[layer], 3)
        assert _rel(a, g[f"a_l{layer}_f64"]).max() < 1e-6
        assert _rel(R, g[f"R_l{layer}_f64"]).max() < 1e-6


def test_arch_b_matches_reference_fixture(golden_dir):
    """The reference's 3-second GTZAN model (cpf.py:410-412: one conv per block, no BatchNorm, 128 x 128) under its own rule
    map LRP_NAME_MAP_GTZAN (constants.py:27-38), split at two of the layers cpf.py:141 uses."""
    g, net = _load(golden_dir, "archB")
    x = synth.synth_logmel(int(g["N"]), 128, 128, int(g["x_seed"]))
    o = lrp_ref.lrp_pass(net, x, LRP_NAME_MAP_GTZAN, lrp_ref.output_modifier(6))
    assert _rel(o["R_input"], g["Rin_c6_f64"]).max() < 1e-6
    np.testing.assert_allclose(o["logits"].numpy(), g["logits_f64"], rtol=1e-5, atol=1e-7)
    for layer, d, hw in ((7, 64, 32), (13, 128, 8)):
        a, R = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_GTZAN, net.features[layer], 6)
        assert a.shape == (3, d, hw, hw)
        assert _rel(a, g[f"a_l{layer}_f64"]).max() < 1e-6
        assert _rel(R, g[f"R_l{layer}_f64"]).max() < 1e-6
    assert max(float(g[k].max()) for k in g.files if k.startswith("noise_")) < TOL


def test_arch_b_early_split_layers_match_reference_fixture(golden_dir):
    """features[1], [4] (d = 32) and [10] (d = 64), the remaining split layers of cpf.py:141 (early maps stored at stride 4)."""
    g, net = _load(golden_dir, "archB_early")
    x = synth.synth_logmel(3, 128, 128, int(g["x_seed"]))[:int(g["N"])]
    for layer, st, d, hw in ((1, 4, 32, 128), (4, 4, 32, 64), (10, 1, 64, 16)):
        a, R = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_GTZAN, net.features[layer], 6)
        assert a.shape == (2, d, hw, hw)
        assert _rel(a[..., ::st, ::st], g[f"a_l{layer}_f64"]).max() < 1e-6
        assert _rel(R[..., ::st, ::st], g[f"R_l{layer}_f64"]).max() < 1e-6


def test_cfg2_full_resolution_maps_match_reference_fixture(golden_dir):
    g, net = _load(golden_dir, "cfg2_full")
    x = synth.synth_logmel(int(g["N"]), 128, 256, int(g["x_seed"]))[:1]       # one sample keeps the CPU suite short
    a, R = lrp_ref.get_intermediate(net, x, lrp_name_map_6s(), net.features[33], 3)
    assert a.shape == (1, 256, 8, 8)
    assert _rel(a, g["a_l33_f64"][:1]).max() < 1e-6
    assert _rel(R, g["R_l33_f64"][:1]).max() < 1e-6


@pytest.mark.parametrize("tag", ["perm", "orth"])
def test_projection_model_matches_reference_fixture(golden_dir, tag):
    """HeatmapGenerator of the reference (explainer.py:15-177) vs the oracle pushing K+1 clones through the product's
    ProjectionModel container."""
    from cxai.model.modify_model import ProjectionModel
    from cxai.xai.explain.explainer import get_class_composite
    g, net = _load(golden_dir, "heat_toy")
    K, layer_idx, d, N = int(g["K"]), int(g["layer_idx"]), int(g["d"]), int(g["N"])
    x = synth.synth_logmel(N, 64, 64, int(g["x_seed"]))
    U = synth.signed_permutation(d, 5) if tag == "perm" else synth.random_orthogonal(d, 6)
    pm = ProjectionModel(net, layer_idx, U.double(), K, case="toy")
    comp = get_class_composite(LRP_NAME_MAP_TOY, K)
    o = lrp_ref.lrp_pass(pm, x.repeat_interleave(K + 1, dim=0), comp.name_map, lrp_ref.output_modifier(0))
    H = o["R_input"].view(N, K + 1, 64, 64)
    assert _rel(H[:, 0], g[f"{tag}_standard_heatmaps"][:, 0]).max() < 1e-6
    mask = g[f"{tag}_mask"]
    sub = np.take_along_axis(H[:, 1:].numpy(), mask[:, :, None, None], axis=1)      # sorted like explainer.py:151-176
    tol = 1e-6 if tag == "perm" else 0.05            # a general U leaves the single concept maps defined to ~1e-2 only
    for j in range(K):
        assert _rel(sub[:, j], g[f"{tag}_subspace_heatmaps"][:, j]).max() < tol
    assert _rel(H[:, 1:].sum(1), g[f"{tag}_subspace_heatmaps"].sum(1)).max() < 1e-5
    # compute_subspace_relevances (explainer.py:206-242) on the maps of the same model
    a, R = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_TOY, net.features[layer_idx], 0)
    av, cv = a.flatten(2).transpose(1, 2), (R / (a + 1e-7)).flatten(2).transpose(1, 2)
    np.testing.assert_allclose(drsa_ref.subspace_relevances(av, cv, U.double(), K).numpy(), g[f"{tag}_Rk"],
                               rtol=2e-4, atol=1e-6 * float(np.abs(g[f"{tag}_Rk"]).max()))


@pytest.mark.parametrize("tag", ["perm", "orth"])
def test_projection_model_arch_b_matches_reference_fixture(golden_dir, tag):
    """The same on arch B (cpf.py:410-412) split at features[13] (d = 128), rule map LRP_NAME_MAP_GTZAN, class 'rock'."""
    from cxai.model.modify_model import ProjectionModel
    from cxai.xai.explain.explainer import get_class_composite
    g, net = _load(golden_dir, "heat_archB")
    K, layer_idx, d, N = int(g["K"]), int(g["layer_idx"]), int(g["d"]), int(g["N"])
    assert (K, layer_idx, d) == (4, 13, 128)
    x = synth.synth_logmel(N, 128, 128, int(g["x_seed"]))
    U = synth.signed_permutation(d, 5) if tag == "perm" else synth.random_orthogonal(d, 6)
    pm = ProjectionModel(net, layer_idx, U.double(), K, case="gtzan")
    comp = get_class_composite(LRP_NAME_MAP_GTZAN, K)
    o = lrp_ref.lrp_pass(pm, x.repeat_interleave(K + 1, dim=0), comp.name_map, lrp_ref.output_modifier(6))
    H = o["R_input"].view(N, K + 1, 128, 128)
    assert _rel(H[:, 0], g[f"{tag}_standard_heatmaps"][:, 0]).max() < max(1e-6, 3 * float(g[f"noise_{tag}_standard_heatmaps"].max()))
    mask = g[f"{tag}_mask"]
    sub = np.take_along_axis(H[:, 1:].numpy(), mask[:, :, None, None], axis=1)
    tol = 1e-6 if tag == "perm" else 0.05
    for j in range(K):
        assert _rel(sub[:, j], g[f"{tag}_subspace_heatmaps"][:, j]).max() < tol
    assert _rel(H[:, 1:].sum(1), g[f"{tag}_subspace_heatmaps"].sum(1)).max() < 1e-4


def _prep_inputs(g):
    gen = torch.Generator().manual_seed(int(g["seed"]))
    B, d, H, W = (int(g[k]) for k in ("B", "d", "H", "W"))
    amap = torch.relu(torch.randn(B, d, H, W, generator=gen))
    Rmap = torch.randn(B, d, H, W, generator=gen) * (amap > 0)
    return amap, Rmap


def test_preprocessing_helpers_match_reference_fixture(golden_dir):
    """sample_spatial_locations / get_vectors_from_maps / compute_context_vectors / normalize_vectors of the reference
    (preprocessing.py:179-256, run unmodified by oracle/gen_golden_lrp.py `prep`) against the oracle's restatements and against
    the host-side parts of the product (numpy RNG call order, the reference's scrambled row layout)."""
    from cxai.xai.drsa import preprocessing as pp
    g = np.load(os.path.join(golden_dir, "lrp_prep.npz"))
    amap, Rmap = _prep_inputs(g)
    B, H, W, L = (int(g[k]) for k in ("B", "H", "W", "L"))
    np.random.seed(int(g["np_seed"]))
    idcs = pp.sample_spatial_locations(B, (H, W), L)                         # product, host only
    np.testing.assert_array_equal(idcs, g["idcs"])
    assert idcs.dtype == g["idcs"].dtype
    np.testing.assert_array_equal(pp.get_vectors_from_maps(amap, idcs, layout="reference").numpy(), g["va"])
    # oracle restatements
    va, vr = drsa_ref.vectors_from_maps_ref(amap, idcs), drsa_ref.vectors_from_maps_ref(Rmap, idcs)
    np.testing.assert_array_equal(va.numpy(), g["va"])
    np.testing.assert_array_equal(vr.numpy(), g["vr"])
    c = drsa_ref.compute_context_vectors(va, vr)
    np.testing.assert_array_equal(c.numpy(), g["c"])
    np.testing.assert_array_equal(drsa_ref.compute_context_vectors(amap, Rmap).numpy(), g["c_maps"])
    np.testing.assert_allclose(drsa_ref.normalize_vectors(va).numpy(), g["na"], rtol=1e-6)
    np.testing.assert_allclose(drsa_ref.normalize_vectors(c).numpy(), g["nc"], rtol=1e-6)
    # the corrected layout holds the same vectors, row (b, l) = position idcs[b, l] of sample b
    fixed = pp.get_vectors_from_maps(amap, idcs, layout="fixed")
    want = torch.stack([amap[b].flatten(1)[:, idcs[b, l]] for b in range(B) for l in range(L)])
    np.testing.assert_array_equal(fixed.numpy(), want.numpy())


def test_mini_zennit_gamma_collapses_and_restores_parameters():
    """The restated hooks leave the model untouched after the context and the 5-pass Gamma equals the one-pass form for
    non-negative input (what the CUDA kernels rely on)."""
    from oracle.mini_zennit.rules import Gamma, Epsilon
    from oracle.mini_zennit.composites import NameMapComposite
    from oracle.mini_zennit.attribution import Gradient
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3, padding=1), torch.nn.BatchNorm2d(5), torch.nn.ReLU(),
                              torch.nn.Flatten(), torch.nn.Linear(5 * 6 * 6, 4)).double().eval()
    synth.randomize_bn(net, 2)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.rand(2, 3, 6, 6, dtype=torch.float64)
    from oracle.mini_zennit.canonizers import SequentialMergeBatchNorm
    comp = NameMapComposite([(["0"], Gamma(gamma=0.3, stabilizer=1e-7)), (["4"], Epsilon(epsilon=1e-7))],
                            canonizers=[SequentialMergeBatchNorm()])
    with Gradient(net, comp) as attributor:
        out, R = attributor(x, lambda o: o * torch.eye(4, dtype=o.dtype)[1])
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert net[1].eps == 1e-5
    # closed form
    bn = net[1]
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = net[0].weight * scale[:, None, None, None]; b = (net[0].bias - bn.running_mean) * scale + bn.bias
    z = torch.nn.functional.conv2d(x, w, b, padding=1)
    h = z.clamp(min=0).flatten(1)
    logits = net[4](h)
    np.testing.assert_allclose(out.detach().numpy(), logits.detach().numpy(), rtol=1e-10)
    Rl = torch.zeros_like(logits); Rl[:, 1] = logits[:, 1]
    Rh = h * ((Rl / lrp_ref.stabilize(logits, 1e-7)) @ net[4].weight)
    Rz = Rh.view_as(z) * (z > 0)
    wm, bm = w + 0.3 * w.clamp(min=0), b + 0.3 * b.clamp(min=0)
    zp = torch.nn.functional.conv2d(x, wm, bm, padding=1)
    want = x * torch.nn.functional.conv_transpose2d(Rz / lrp_ref.stabilize(zp, 1e-7), wm, padding=1)
    np.testing.assert_allclose(R.detach().numpy(), want.detach().numpy(), rtol=1e-8, atol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference/cxai"), reason="reference tree only exists in the build container")
def test_fixture_generator_reproduces_committed_fixture(golden_dir, tmp_path):
    """Re-runs the reference's code (toy case) and compares with the committed file: the fixtures are what the
    generator script says they are."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, runpy; sys.argv = ['gen', 'archA_small']; import oracle.gen_golden_lrp as G; "
            f"G.GOLD = {str(tmp_path)!r}; import numpy as np; out = G.CASES['archA_small'](); "
            f"np.savez({str(tmp_path / 'x.npz')!r}, **out)")
    env = dict(os.environ, PYTHONPATH=root)
    subprocess.run([sys.executable, "-c", code], check=True, cwd=root, env=env, capture_output=True, timeout=600)
    new, old = np.load(tmp_path / "x.npz"), np.load(os.path.join(golden_dir, "lrp_archA_small.npz"))
    for k in old.files:
        if old[k].dtype.kind == "f" and not k.startswith("noise_"):
            np.testing.assert_allclose(new[k], old[k], rtol=1e-5, atol=1e-9, err_msg=k)
