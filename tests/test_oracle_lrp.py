"""Self-checks of the LRP oracle (CPU).  Parity with zennit itself is UNPINNED (package unavailable,
reference has no golden vectors): these tests pin the restated semantics through the invariants of
SURVEY section 4 / 8(c)."""
import numpy as np
import torch
import torch.nn as nn

from oracle import lrp_ref
from cxai.utils.constants import LRP_NAME_MAP_TOY
from cxai.xai.explain.rules import Epsilon, Gamma, ZPlus


def test_stabilize_zero_counts_as_positive():
    x = torch.tensor([-2.0, 0.0, 3.0])
    np.testing.assert_allclose(lrp_ref.stabilize(x, 0.5).numpy(), [-2.5, 0.5, 3.5])


def test_gamma_general_form_collapses_for_nonnegative_input():
    """For x >= 0 and R supported on z > 0 the 5-pass Gamma equals x * W'^T (R / stab(W'x+b'))."""
    torch.manual_seed(0)
    conv = nn.Conv2d(5, 7, 3, padding=1).double()
    x = torch.rand(3, 5, 9, 11, dtype=torch.float64)
    x[x < 0.3] = 0
    z = conv(x).detach()
    R = torch.randn_like(z) * (z > 0)
    gamma, eps = 0.3, 1e-7
    general = lrp_ref.rule_backward(conv, "gamma", x, R, eps, gamma)
    w = conv.weight.detach(); b = conv.bias.detach()
    wm, bm = w + gamma * w.clamp(min=0), b + gamma * b.clamp(min=0)
    zp = torch.nn.functional.conv2d(x, wm, bm, padding=1)
    s = R / lrp_ref.stabilize(zp, eps)
    collapsed = x * torch.nn.functional.conv_transpose2d(s, wm, padding=1)
    np.testing.assert_allclose(general.numpy(), collapsed.numpy(), rtol=1e-10, atol=1e-12)
    # ZPlus is the gamma -> infinity limit
    zplus = lrp_ref.rule_backward(conv, "zplus", x, R, eps)
    big = lrp_ref.rule_backward(conv, "gamma", x, R, eps, 1e9)
    np.testing.assert_allclose(zplus.numpy(), big.numpy(), rtol=1e-3, atol=1e-6)


def test_epsilon_is_gradient_times_input_on_biasfree_relu_net():
    torch.manual_seed(1)
    net = nn.Sequential(nn.Linear(6, 8, bias=False), nn.ReLU(), nn.Linear(8, 3, bias=False)).double()
    x = torch.randn(4, 6, dtype=torch.float64, requires_grad=True)
    out = net(x)
    gi, = torch.autograd.grad(out[:, 1].sum(), x)
    R = torch.zeros_like(out); R[:, 1] = out[:, 1].detach()
    h = net[0](x).clamp(min=0).detach()
    R = lrp_ref.rule_backward(net[2], "epsilon", h, R, 1e-12)
    R = R * (h > 0)
    R = lrp_ref.rule_backward(net[0], "epsilon", x.detach(), R, 1e-12)
    np.testing.assert_allclose(R.numpy(), (x * gi).detach().numpy(), rtol=1e-6, atol=1e-9)


def test_conservation_and_split_layer_on_toy_model():
    net = lrp_ref.toy_model(seed=0, last=16)
    for m in net.modules():                       # bias-free => relevance is conserved layer to layer
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            m.bias.data.zero_()
    x = lrp_ref.synth_logmel(4, 64, 64, 3)
    layer = net.features[13]
    o = lrp_ref.lrp_pass(net, x, LRP_NAME_MAP_TOY, lrp_ref.output_modifier(1), split_module=layer)
    logit = o["logits"][:, 1]
    # conserved up to the stabiliser (1e-7 against |z| ~ 1e-2 .. 1e-3)
    np.testing.assert_allclose(o["R_split"].sum(dim=(1, 2, 3)).numpy(), logit.numpy(), rtol=5e-4)
    # down to the input with WSquare on the first conv (Flat turns the bias into ones, which absorbs relevance)
    from cxai.xai.explain.rules import WSquare
    nm = [(["features.0"], WSquare(stabilizer=1e-7))] + LRP_NAME_MAP_TOY[1:]
    o3 = lrp_ref.lrp_pass(net, x, nm, lrp_ref.output_modifier(1))
    np.testing.assert_allclose(o3["R_input"].sum(dim=(1, 2, 3)).numpy(), logit.numpy(), rtol=5e-3)
    a, R = o["a_split"], o["R_split"]
    assert a.shape == (4, 16, 4, 4) and float(a.min()) >= 0
    # Gamma layers return x * (...): relevance INTO the first Gamma conv vanishes where its input is 0
    o2 = lrp_ref.lrp_pass(net, x, LRP_NAME_MAP_TOY, lrp_ref.output_modifier(1), split_module=net.features[2])
    assert float(o2["R_split"][o2["a_split"] == 0].abs().max()) == 0.0


def test_get_intermediate_minibatching_is_transparent():
    net = lrp_ref.toy_model(seed=2, last=16)
    x = lrp_ref.synth_logmel(5, 64, 64, 4)
    a1, r1 = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_TOY, net.features[13], 0, attr_batch_size=2)
    a2, r2 = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_TOY, net.features[13], 0, attr_batch_size=64)
    np.testing.assert_allclose(a1.numpy(), a2.numpy(), rtol=1e-12)
    np.testing.assert_allclose(r1.numpy(), r2.numpy(), rtol=1e-10, atol=1e-14)


def test_batchnorm_merge_matches_eval_forward():
    net = lrp_ref.genre_model(seed=0, last=32, input_size=(32, 64))
    x = lrp_ref.synth_logmel(2, 32, 64, 5)
    from cxai.utils.constants import lrp_name_map_6s
    o = lrp_ref.lrp_pass(net, x, lrp_name_map_6s(), lrp_ref.output_modifier(3), split_module=net.features[33])
    with torch.no_grad():
        ref_logits = net.double()(x.double())
    np.testing.assert_allclose(o["logits"].numpy(), ref_logits.numpy(), rtol=1e-4, atol=1e-7)
    assert o["a_split"].shape == (2, 32, 2, 2)


def test_projection_model_oracle_invariants():
    """ProjectionModel + SubspaceHook (modify_model.py:4-123, attribute.py:12-67, explainer.py:186-203) in the oracle:
    the concept heatmaps add up to the standard heatmap (every layer below the filter is linear in the relevance),
    clone 0 equals the plain LRP pass up to the stabilisers of the two projection layers, and the projections do not
    change the logits for an orthogonal U."""
    import torch
    from oracle import drsa_ref
    from cxai.utils.constants import LRP_NAME_MAP_TOY
    from cxai.model.modify_model import ProjectionModel
    from cxai.xai.explain.explainer import get_class_composite
    net = lrp_ref.toy_model(seed=0, last=64)
    x = lrp_ref.synth_logmel(3, 64, 64, 5)
    K = 4
    U = torch.linalg.qr(torch.randn(16, 16, dtype=torch.float64, generator=torch.Generator().manual_seed(3)))[0]
    pm = ProjectionModel(net, 10, U, K, case="toy")
    comp = get_class_composite(LRP_NAME_MAP_TOY, K)
    o = lrp_ref.lrp_pass(pm, x.repeat_interleave(K + 1, dim=0), comp.name_map, lrp_ref.output_modifier(1))
    Hm = o["R_input"].view(3, K + 1, 64, 64)
    plain = lrp_ref.lrp_pass(net, x, LRP_NAME_MAP_TOY, lrp_ref.output_modifier(1))
    assert float((Hm[:, 1:].sum(1) - Hm[:, 0]).abs().max() / Hm[:, 0].abs().max()) < 1e-12
    assert float((Hm[:, 0] - plain["R_input"].view(3, 64, 64)).abs().max() / plain["R_input"].abs().max()) < 1e-3
    assert float((o["logits"][::K + 1] - plain["logits"]).abs().max()) < 1e-12
    # the torch forward of the ProjectionModel itself agrees as well (InvProjection handles non-square maps)
    with torch.no_grad():
        assert float((pm.double()(x.double()) - net.double()(x.double())).abs().max()) < 1e-10
    net.float()
