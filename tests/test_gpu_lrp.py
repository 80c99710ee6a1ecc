"""GPU parity tests of stage 1 (forward + LRP pass, context extraction) against the CPU oracle
(oracle/lrp_ref.py, the general multi-pass restatement of zennit's rules; parity with zennit itself is
unpinned, see that file's header).  Metric: norm-wise relative error per sample <= 1e-4."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import lrp_ref, drsa_ref

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _rel_per_sample(got, want):
    got, want = got.double().cpu().flatten(1), want.double().cpu().flatten(1)
    return float(((got - want).norm(dim=1) / want.norm(dim=1).clamp(min=1e-30)).max())


def _L():
    from drsa_audio_b200 import _lib
    return _lib


@pytest.mark.parametrize("N,Cin,Cout,H,W", [(2, 1, 8, 64, 64), (3, 8, 16, 17, 33), (2, 64, 100, 16, 40), (1, 100, 128, 8, 8),
                                            (2, 5, 7, 3, 5)])
def test_conv3x3_forward_and_rule_backward(N, Cin, Cout, H, W):
    L = _L(); lib = L.lib()
    g = torch.Generator().manual_seed(Cin * Cout)
    x = torch.rand(N, Cin, H, W, generator=g); x[x < 0.2] = 0
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = 0.1 * torch.randn(Cout, generator=g)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    s = torch.cuda.current_stream().cuda_stream
    for relu in (0, 1):
        y = torch.empty(N, Cout, H, W, device="cuda")
        L.check(lib.lrp_conv3x3_forward(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), N, Cin, Cout, H, W, relu, y.data_ptr(), s))
        want = F.conv2d(x.double(), w.double(), b.double(), padding=1)
        if relu:
            want = want.clamp(min=0)
        assert _rel_per_sample(y, want) < 1e-5
    # Gamma-style backward with modified weights vs the general 5-pass oracle
    conv = torch.nn.Conv2d(Cin, Cout, 3, padding=1).double()
    conv.weight.data.copy_(w); conv.bias.data.copy_(b)
    z = conv(x.double()).detach()
    Rout = (torch.randn(z.shape, generator=g).double() * (z > 0)).float()
    gamma, eps = 0.3, 1e-7
    want = lrp_ref.rule_backward(conv, "gamma", x.double(), Rout.double(), eps, gamma)
    wm = (w + gamma * w.clamp(min=0)).cuda().contiguous(); bm = (b + gamma * b.clamp(min=0)).cuda().contiguous()
    wt = torch.empty(Cin, Cout, 3, 3, device="cuda")
    L.check(lib.lrp_conv3x3_flip_weights(wm.data_ptr(), Cout, Cin, wt.data_ptr(), s))
    sbuf = torch.empty(N, Cout, H, W, device="cuda"); Rin = torch.empty(N, Cin, H, W, device="cuda")
    Rd = Rout.cuda()
    L.check(lib.lrp_conv3x3_backward(xd.data_ptr(), wm.data_ptr(), wt.data_ptr(), bm.data_ptr(), Rd.data_ptr(), N, Cin,
                                     Cout, H, W, eps, 0, sbuf.data_ptr(), Rin.data_ptr(), s))
    assert _rel_per_sample(Rin, want) < TOL
    # WSquare (input replaced by ones, no input factor)
    want = lrp_ref.rule_backward(conv, "wsquare", x.double(), Rout.double(), eps)
    w2 = (w * w).cuda().contiguous(); b2 = (b * b).cuda().contiguous()
    L.check(lib.lrp_conv3x3_flip_weights(w2.data_ptr(), Cout, Cin, wt.data_ptr(), s))
    L.check(lib.lrp_conv3x3_backward(None, w2.data_ptr(), wt.data_ptr(), b2.data_ptr(), Rd.data_ptr(), N, Cin, Cout, H,
                                     W, eps, 1, sbuf.data_ptr(), Rin.data_ptr(), s))
    assert _rel_per_sample(Rin, want) < TOL


def test_maxpool_and_dense_blocks():
    L = _L(); lib = L.lib()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 8, 12, generator=g); x[0, 0, :2, :4] = 1.5            # ties: first maximum wins
    for kh, kw in ((2, 2), (2, 4)):
        xd = x.cuda()
        Ho, Wo = 8 // kh, 12 // kw
        y = torch.empty(3, 5, Ho, Wo, device="cuda"); am = torch.empty(3, 5, Ho, Wo, dtype=torch.int32, device="cuda")
        L.check(lib.lrp_maxpool_forward(xd.data_ptr(), 15, 8, 12, kh, kw, y.data_ptr(), am.data_ptr(), s))
        want, idx = F.max_pool2d(x, (kh, kw), return_indices=True)
        np.testing.assert_array_equal(y.cpu().numpy(), want.numpy())
        np.testing.assert_array_equal(am.cpu().numpy(), idx.numpy().astype(np.int32))
        Rout = torch.randn(3, 5, Ho, Wo, generator=g)
        Rin = torch.empty(3, 5, 8, 12, device="cuda")
        Rd = Rout.cuda()
        L.check(lib.lrp_maxpool_backward(Rd.data_ptr(), am.data_ptr(), 15, 8, 12, kh, kw, Rin.data_ptr(), s))
        xi = x.clone().requires_grad_(True)
        gr, = torch.autograd.grad(F.max_pool2d(xi, (kh, kw)), xi, Rout)
        np.testing.assert_array_equal(Rin.cpu().numpy(), gr.numpy())
    # dense forward + epsilon rule (seeded: R / z amplifies fp32 rounding without bound when some z is ~0)
    torch.manual_seed(3)
    lin = torch.nn.Linear(37, 11).double()
    xv = torch.rand(9, 37, generator=g)
    y = torch.empty(9, 11, device="cuda")
    wd, bd = lin.weight.detach().float().cuda().contiguous(), lin.bias.detach().float().cuda().contiguous()
    xvd = xv.cuda()
    L.check(lib.lrp_dense_forward(xvd.data_ptr(), wd.data_ptr(), bd.data_ptr(), 9, 37, 11, 1, y.data_ptr(), s))
    np.testing.assert_allclose(y.cpu().numpy(), lin(xv.double()).clamp(min=0).detach().numpy(), rtol=1e-5, atol=1e-6)
    Rout = torch.randn(9, 11, generator=g)
    want = lrp_ref.rule_backward(lin, "epsilon", xv.double(), Rout.double(), 1e-7)
    sbuf = torch.empty(9, 11, device="cuda"); Rin = torch.empty(9, 37, device="cuda")
    Rd = Rout.cuda()
    L.check(lib.lrp_dense_epsilon_backward(xvd.data_ptr(), wd.data_ptr(), bd.data_ptr(), Rd.data_ptr(), 9, 37, 11,
                                           1e-7, sbuf.data_ptr(), Rin.data_ptr(), s))
    assert _rel_per_sample(Rin, want) < TOL


def test_get_intermediate_toy_cfg1():
    """cfg 1: toy CNN on 64x64, LRP_NAME_MAP_TOY, split at features[13] -> maps [N,64,4,4]."""
    from cxai.utils.constants import LRP_NAME_MAP_TOY
    from cxai.xai.explain.rules import NameMapComposite
    from cxai.xai.drsa.preprocessing import get_intermediate
    net = lrp_ref.toy_model(seed=0, last=64)
    x = lrp_ref.synth_logmel(70, 64, 64, 20261)               # 2 minibatches (64 + 6)
    comp = NameMapComposite(LRP_NAME_MAP_TOY)
    for cls, onehot in ((0, False), (1, True)):
        a, R = get_intermediate(net, x, comp, net.features[13], cls, one_hot_encoded=onehot)
        aw, Rw = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_TOY, net.features[13], cls, one_hot_encoded=onehot)
        assert a.shape == (70, 64, 4, 4)
        assert _rel_per_sample(a, aw) < 1e-5
        assert _rel_per_sample(R, Rw) < TOL
        assert float(a.min()) >= 0


def test_compute_relevances_full_depth_and_bn_model():
    """arch-A-style BatchNorm model (reduced resolution), production name map, relevance at the input
    (the 'LRP relevance maps' parity surface) and maps at layers 19 / 26 / 33."""
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.xai.drsa.preprocessing import get_intermediate
    net = lrp_ref.genre_model(seed=0, last=64, input_size=(32, 64))
    x = lrp_ref.synth_logmel(6, 32, 64, 20262)
    nm = lrp_name_map_6s()
    comp = NameMapComposite(nm, canonizers=[SequentialMergeBatchNorm()])
    Rin = compute_relevances(net, x, comp, class_idx=3)
    want = lrp_ref.lrp_pass(net, x, nm, lrp_ref.output_modifier(3))
    assert Rin.shape == x.shape
    assert _rel_per_sample(Rin, want["R_input"]) < TOL
    for li in (19, 26, 33):
        a, R = get_intermediate(net, x, comp, net.features[li], 3)
        aw, Rw = lrp_ref.get_intermediate(net, x, nm, net.features[li], 3)
        assert _rel_per_sample(a, aw) < 1e-5, li
        assert _rel_per_sample(R, Rw) < TOL, li
    # balanced batch of all classes (attribute.py:148-158)
    net2 = lrp_ref.toy_model(seed=3, last=16)
    from cxai.utils.constants import LRP_NAME_MAP_TOY
    x2 = lrp_ref.synth_logmel(4, 64, 64, 9)
    R2 = compute_relevances(net2, x2, NameMapComposite(LRP_NAME_MAP_TOY), num_classes=2)
    w2 = lrp_ref.lrp_pass(net2, x2, LRP_NAME_MAP_TOY, lrp_ref.output_modifier(None, 2))
    assert _rel_per_sample(R2, w2["R_input"]) < TOL


def test_end_to_end_cfg1_pipeline():
    """BASELINE cfg 1 end to end: toy CNN -> LRP context at features[13] (d = 64, P = 16) -> normalise ->
    DRSA K = 4; the CUDA pipeline must follow the CPU oracle pipeline (objective 1e-4, angles 1e-3)."""
    from cxai.utils.constants import LRP_NAME_MAP_TOY
    from cxai.xai.explain.rules import NameMapComposite
    from cxai.xai.drsa import preprocessing as pp
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    net = lrp_ref.toy_model(seed=0, last=64)
    N, K, steps = 250, 4, 40
    x = lrp_ref.synth_logmel(N, 64, 64, 20261)
    a, R = pp.get_intermediate(net, x, NameMapComposite(LRP_NAME_MAP_TOY), net.features[13], 0)
    act, ctx = pp.gather_context_pairs(a, R, None, normalize=True)
    aw, Rw = lrp_ref.get_intermediate(net, x, LRP_NAME_MAP_TOY, net.features[13], 0)
    av = drsa_ref.vectors_from_maps_all(aw.float()); rv = drsa_ref.vectors_from_maps_all(Rw.float())
    A = drsa_ref.normalize_vectors(av); C = drsa_ref.normalize_vectors(drsa_ref.compute_context_vectors(av, rv))
    assert act.shape == (N * 16, 64)
    assert _rel_per_sample(act, A) < 1e-4
    U0 = drsa_ref.synth_U0(64, seed=5)
    objs_ref, U_ref = drsa_ref.run_autograd(A, C, U0, K, steps)
    opt = SubspaceOptimizer(U0, act, ctx, None, num_concepts=K)
    opt.run(steps=steps, save=False)
    rel = np.max(np.abs(opt.obj_history - objs_ref) / np.abs(objs_ref))
    ang = drsa_ref.principal_angle(opt.U.cpu(), U_ref, K)
    print(f"cfg1 e2e: objective rel err {rel:.2e}, angle {ang:.2e}")
    assert rel < 1e-4 and ang < 1e-3


def _assert_relevance_close_up_to_pool_ties(a, R1, R2, pool, tol=1e-4, tie=1e-4):
    """R1 ~ R2 except inside pooling windows whose two largest activations tie within `tie` (relative): max-pool
    routing is discontinuous there, so two correct implementations that differ by rounding may pick different
    winners.  Such windows must be rare and must conserve the routed relevance."""
    a, R1, R2 = a.double().cpu(), R1.double().cpu(), R2.double().cpu()
    kh, kw = pool
    B, C, H, W = a.shape
    win = lambda t: t.reshape(B, C, H // kh, kh, W // kw, kw).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // kh, W // kw, kh * kw)
    aw, r1, r2 = win(a), win(R1), win(R2)
    top2 = aw.topk(2, dim=-1).values
    tied = (top2[..., 0] - top2[..., 1]) <= tie * top2[..., 0].abs().clamp(min=1e-30)
    scale = R2.abs().amax(dim=(1, 2, 3), keepdim=True).unsqueeze(-1)
    bad = ((r1 - r2).abs() > tol * scale).any(dim=-1)
    assert not bool((bad & ~tied).any()), "relevance differs outside tied pooling windows"
    assert float(bad.double().mean()) < 1e-3
    # relevance routed through a tied window is the same, only its position differs
    np.testing.assert_allclose(r1.sum(-1)[bad].numpy(), r2.sum(-1)[bad].numpy(), rtol=1e-3, atol=1e-9)
    ok = ~bad
    err = ((r1 - r2) * ok.unsqueeze(-1)).flatten(1).norm(dim=1) / r2.flatten(1).norm(dim=1)
    assert float(err.max()) < tol


def _to_nhwc_split(x, Cp):
    """[B,C,H,W] fp32 -> NHWC hi/lo fp16 planes with channels zero-padded to Cp."""
    B, C, H, W = x.shape
    t = torch.zeros(B, H, W, Cp)
    t[..., :C] = x.permute(0, 2, 3, 1)
    hi = t.half()
    lo = (t - hi.float()).half()
    return hi.cuda().contiguous(), lo.cuda().contiguous()


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(3, 64, 64, 16, 64), (2, 64, 64, 4, 256), (5, 100, 128, 8, 8), (3, 128, 256, 4, 4),
                                            (2, 256, 256, 8, 8), (70, 64, 100, 2, 2), (1, 64, 64, 128, 256)])
def test_tc_conv3x3_forward_matches_fp64(B, Cin, Cout, H, W):
    """tcgen05 implicit-GEMM convolution (TMA im2col, hi/lo split operands) vs an fp64 convolution."""
    L = _L(); lib = L.lib()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(B * Cin + Cout)
    x = torch.rand(B, Cin, H, W, generator=g); x[x < 0.3] = 0
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = 0.1 * torch.randn(Cout, generator=g)
    Cin_p, Cout_p = (Cin + 63) // 64 * 64, (Cout + 63) // 64 * 64
    assert lib.lrp_tc_conv3x3_supported(B, Cin_p, Cout_p, H, W) == 0
    xh, xl = _to_nhwc_split(x, Cin_p)
    wt = torch.zeros(9, Cout_p, Cin_p)
    wt[:, :Cout, :Cin] = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin)
    wtd = wt.cuda()
    wh = torch.empty(9, Cout_p, Cin_p, dtype=torch.float16, device="cuda"); wl = torch.empty_like(wh)
    L.check(lib.lrp_tc_split_f16(wtd.data_ptr(), wtd.numel(), wh.data_ptr(), wl.data_ptr(), s))
    bias = torch.zeros(Cout_p); bias[:Cout] = b
    bd = bias.cuda()
    yh = torch.empty(B, H, W, Cout_p, dtype=torch.float16, device="cuda"); yl = torch.empty_like(yh)
    yn = torch.full((B, Cout, H, W), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check(lib.lrp_tc_conv3x3_forward(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), bd.data_ptr(), B, H, W,
                                       Cin_p, Cout_p, Cout, 1, yh.data_ptr(), yl.data_ptr(), yn.data_ptr(), err.data_ptr(), s))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1).clamp(min=0)
    assert _rel_per_sample(yn, want) < 2e-5          # hi/lo fp16 split: ~2^-22 per operand, lo*lo dropped
    y2 = (yh.float() + yl.float()).cpu()[..., :Cout].permute(0, 3, 1, 2)
    assert _rel_per_sample(y2, want) < 2e-5
    assert float((yh.float() + yl.float())[..., Cout:].abs().max()) == 0.0 if Cout_p > Cout else True


def test_tc_first_conv_pool_and_layout_roundtrip():
    L = _L(); lib = L.lib()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(4)
    B, H, W, Cout = 3, 16, 32, 64
    x = lrp_ref.synth_logmel(B, H, W, 8)
    w = torch.randn(Cout, 1, 3, 3, generator=g) / 3; b = 0.1 * torch.randn(Cout, generator=g)
    xd, wd, bd = x.cuda(), w.reshape(Cout, 9).cuda().contiguous(), b.cuda()
    yh = torch.empty(B, H, W, Cout, dtype=torch.float16, device="cuda"); yl = torch.empty_like(yh)
    L.check(lib.lrp_tc_conv3x3_first(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), B, H, W, Cout, Cout, 1, yh.data_ptr(),
                                     yl.data_ptr(), s))
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1).clamp(min=0)
    got = (yh.float() + yl.float()).cpu().permute(0, 3, 1, 2)
    assert _rel_per_sample(got, want) < 2e-6
    ph = torch.empty(B, H // 2, W // 4, Cout, dtype=torch.float16, device="cuda"); pl = torch.empty_like(ph)
    L.check(lib.lrp_tc_maxpool(yh.data_ptr(), yl.data_ptr(), B, H, W, Cout, 2, 4, ph.data_ptr(), pl.data_ptr(), None, s))
    out = torch.empty(B, Cout, H // 2, W // 4, device="cuda")
    L.check(lib.lrp_tc_nhwc_to_nchw(ph.data_ptr(), pl.data_ptr(), B, H // 2, W // 4, Cout, Cout, out.data_ptr(), s))
    wantp = F.max_pool2d(got, (2, 4))
    np.testing.assert_allclose(out.cpu().numpy(), wantp.numpy(), rtol=1e-6, atol=1e-7)


def test_tc_rule_backward_kernels_match_oracle():
    """tensor-core ratio + input-multiply passes (the collapsed Gamma rule) vs the general 5-pass oracle."""
    L = _L(); lib = L.lib()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(11)
    B, Cin, Cout, H, W = 3, 64, 100, 8, 16
    Cin_p, Cout_p = 64, 128
    x = torch.rand(B, Cin, H, W, generator=g); x[x < 0.3] = 0
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = 0.1 * torch.randn(Cout, generator=g)
    conv = torch.nn.Conv2d(Cin, Cout, 3, padding=1).double()
    conv.weight.data.copy_(w); conv.bias.data.copy_(b)
    z = conv(x.double()).detach()
    Rout = (torch.randn(z.shape, generator=g).double() * (z > 0)).float()
    gamma, eps = 0.3, 1e-7
    want = lrp_ref.rule_backward(conv, "gamma", x.double(), Rout.double(), eps, gamma)
    wm = w + gamma * w.clamp(min=0); bm = b + gamma * b.clamp(min=0)

    def planes(t):
        td = t.cuda().contiguous()
        hi = torch.empty(td.shape, dtype=torch.float16, device="cuda"); lo = torch.empty_like(hi)
        L.check(lib.lrp_tc_split_f16(td.data_ptr(), td.numel(), hi.data_ptr(), lo.data_ptr(), s))
        return hi, lo
    wt = torch.zeros(9, Cout_p, Cin_p); wt[:, :Cout, :Cin] = wm.permute(2, 3, 0, 1).reshape(9, Cout, Cin)
    tt = torch.zeros(9, Cin_p, Cout_p); tt[:, :Cin, :Cout] = wm.permute(2, 3, 1, 0).reshape(9, Cin, Cout).flip(0)
    mh, ml = planes(wt); th, tl = planes(tt)
    bias = torch.zeros(Cout_p); bias[:Cout] = bm; bd = bias.cuda()
    xh, xl = _to_nhwc_split(x, Cin_p)
    Rn = torch.zeros(B, H, W, Cout_p); Rn[..., :Cout] = Rout.permute(0, 2, 3, 1); Rd = Rn.cuda().contiguous()
    sh = torch.empty(B, H, W, Cout_p, dtype=torch.float16, device="cuda"); sl = torch.empty_like(sh)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check(lib.lrp_tc_conv3x3_ratio(xh.data_ptr(), xl.data_ptr(), mh.data_ptr(), ml.data_ptr(), bd.data_ptr(), Rd.data_ptr(), B, H,
                                     W, Cin_p, Cout_p, eps, None, sh.data_ptr(), sl.data_ptr(), err.data_ptr(), s))
    Rin = torch.empty(B, H, W, Cin_p, device="cuda")
    cmax = torch.zeros(B, device="cuda")
    L.check(lib.lrp_tc_conv3x3_inputmul(sh.data_ptr(), sl.data_ptr(), th.data_ptr(), tl.data_ptr(), xh.data_ptr(), xl.data_ptr(), B,
                                        H, W, Cout_p, Cin_p, None, cmax.data_ptr(), Rin.data_ptr(), err.data_ptr(), s))
    out = torch.empty(B, Cin, H, W, device="cuda")
    L.check(lib.lrp_tc_nhwc_f32_to_nchw(Rin.data_ptr(), B, H, W, Cin_p, Cin, out.data_ptr(), s))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    assert _rel_per_sample(out, want) < TOL
    # cmax = per-sample max |c| with c = conv_transpose(s, w') (the bound handed to the layer below)
    z_mod = F.conv2d(x.double(), wm.double(), bm.double(), padding=1)
    s_want = Rout.double() / (z_mod + torch.where(z_mod >= 0, eps, -eps))
    c_want = F.conv_transpose2d(s_want, wm.double(), padding=1).abs()
    np.testing.assert_allclose(cmax.cpu().numpy(), c_want.amax(dim=(1, 2, 3)).numpy(), rtol=1e-3)
    # relevance far below the fp16 range (it shrinks by orders of magnitude on the way down the network), very
    # different per sample: the per-sample power-of-two scale keeps full accuracy
    fac = torch.tensor([1e-11, 3e-20, 7.0]).view(B, 1, 1, 1)
    Rd2 = (Rn * fac).cuda().contiguous()
    bound = ((s_want * fac).abs().amax(dim=(1, 2, 3)) * 1.5).float().cuda()
    L.check(lib.lrp_tc_conv3x3_ratio(xh.data_ptr(), xl.data_ptr(), mh.data_ptr(), ml.data_ptr(), bd.data_ptr(), Rd2.data_ptr(), B,
                                     H, W, Cin_p, Cout_p, eps, bound.data_ptr(), sh.data_ptr(), sl.data_ptr(), err.data_ptr(), s))
    assert float(sh.float().abs().max()) >= 1024.0                      # planes sit high in the fp16 range
    cmax.zero_()
    L.check(lib.lrp_tc_conv3x3_inputmul(sh.data_ptr(), sl.data_ptr(), th.data_ptr(), tl.data_ptr(), xh.data_ptr(), xl.data_ptr(), B,
                                        H, W, Cout_p, Cin_p, bound.data_ptr(), cmax.data_ptr(), Rin.data_ptr(), err.data_ptr(), s))
    L.check(lib.lrp_tc_nhwc_f32_to_nchw(Rin.data_ptr(), B, H, W, Cin_p, Cin, out.data_ptr(), s))
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    assert _rel_per_sample(out, want * fac.double()) < TOL
    np.testing.assert_allclose(cmax.cpu().numpy(), (c_want * fac.double()).amax(dim=(1, 2, 3)).float().numpy(), rtol=1e-3)
    # a bound that is far too small drives s out of the fp16 range: the kernel must flag it, not return garbage silently
    tiny = (bound * 1e-4).contiguous()
    L.check(lib.lrp_tc_conv3x3_ratio(xh.data_ptr(), xl.data_ptr(), mh.data_ptr(), ml.data_ptr(), bd.data_ptr(), Rd2.data_ptr(), B,
                                     H, W, Cin_p, Cout_p, eps, tiny.data_ptr(), sh.data_ptr(), sl.data_ptr(), err.data_ptr(), s))
    torch.cuda.synchronize()
    assert int(err.item()) == 2
    # per-sample max |R / x| over x > 0 (entry bound below the dense head)
    xr = torch.rand(4, 300, generator=g); xr[xr < 0.4] = 0
    cr = torch.randn(4, 300, generator=g) * torch.tensor([1.0, 1e-9, 50.0, 0.0]).view(4, 1)
    Rr = (xr * cr).cuda(); xrd = xr.cuda(); ob = torch.empty(4, device="cuda")
    L.check(lib.lrp_tc_sample_absmax_ratio(Rr.data_ptr(), xrd.data_ptr(), 4, 300, ob.data_ptr(), s))
    np.testing.assert_allclose(ob.cpu().numpy(), (cr.abs() * (xr > 0)).amax(dim=1).numpy(), rtol=1e-5)
    # NHWC pooling with arg-max capture and relevance routing vs autograd
    a = torch.rand(2, 8, 6, 64, generator=g)
    ah, al = a.half().cuda(), (a - a.half().float()).half().cuda()
    ph = torch.empty(2, 4, 3, 64, dtype=torch.float16, device="cuda"); pl = torch.empty_like(ph)
    am = torch.empty(2, 4, 3, 64, dtype=torch.uint8, device="cuda")
    L.check(lib.lrp_tc_maxpool(ah.data_ptr(), al.data_ptr(), 2, 8, 6, 64, 2, 2, ph.data_ptr(), pl.data_ptr(), am.data_ptr(), s))
    Ro = torch.randn(2, 4, 3, 64, generator=g).cuda()
    Ri = torch.empty(2, 8, 6, 64, device="cuda")
    L.check(lib.lrp_tc_maxpool_backward(Ro.data_ptr(), am.data_ptr(), 2, 8, 6, 64, 2, 2, Ri.data_ptr(), s))
    a_rec = (ah.float() + al.float()).cpu().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gr, = torch.autograd.grad(F.max_pool2d(a_rec, 2), a_rec, Ro.cpu().permute(0, 3, 1, 2))
    np.testing.assert_array_equal(Ri.cpu().permute(0, 3, 1, 2).numpy(), gr.numpy())


def test_tc_prefix_equals_fp32_path_full_resolution():
    """cfg-2 CNN at full 128x256 resolution: the tensor-core stack and the CUDA-core path give the same maps."""
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.explain import lrp_engine
    net = lrp_ref.genre_model(seed=0, last=256, input_size=(128, 256))
    x = lrp_ref.synth_logmel(3, 128, 256, 20262).cuda()
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    plan = lrp_engine._plan(net, comp, x.device)
    from cxai.xai.drsa.preprocessing import get_intermediate
    res = {}
    for use_tc in (True, False):
        plan.use_tc = use_tc
        res[use_tc] = get_intermediate(net, x, comp, net.features[33], 2)
    plan.use_tc = True
    assert plan._tc_stack_ok(x)
    assert res[True][0].shape == (3, 256, 8, 8)
    assert _rel_per_sample(res[True][0], res[False][0]) < 1e-5
    _assert_relevance_close_up_to_pool_ties(res[True][0], res[True][1], res[False][1], (2, 2))


def test_engine_chunk_does_not_change_results(monkeypatch):
    """The engine takes up to ENGINE_CHUNK samples per pass instead of the reference's minibatches of 64
    (preprocessing.py:150-167); samples are independent, so the maps must not depend on the cut."""
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.explain import lrp_engine
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.xai.drsa.preprocessing import get_intermediate
    net = lrp_ref.genre_model(seed=0, last=64, input_size=(32, 64))
    x = lrp_ref.synth_logmel(7, 32, 64, 20263).cuda()
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    assert lrp_engine._engine_chunk(x, 2) >= 7
    whole = get_intermediate(net, x, comp, net.features[26], 1, attr_batch_size=2)
    Rw = compute_relevances(net, x, comp, class_idx=1)
    monkeypatch.setattr(lrp_engine, "ENGINE_CHUNK", 1)
    assert lrp_engine._engine_chunk(x, 2) == 2
    cut = get_intermediate(net, x, comp, net.features[26], 1, attr_batch_size=2)
    monkeypatch.setattr(lrp_engine, "ENGINE_CHUNK", 3)
    Rc = compute_relevances(net, x, comp, class_idx=1)
    # (kernel selection in the dense head may depend on the row count: summation order, not arithmetic, differs)
    assert _rel_per_sample(whole[0], cut[0]) < 1e-6 and _rel_per_sample(whole[1], cut[1]) < 1e-5
    assert _rel_per_sample(Rw, Rc) < 1e-5


@pytest.mark.parametrize("case", ["genre_bn", "toy"])
def test_fused_conv_pool_equals_separate_kernels(case):
    """MaxPool2d fused into the epilogue of the convolution before it (lrp_tc_conv3x3_forward_pool) against the separate
    pooling kernel: activations and relevance at a split layer (pools below the split are fused), and the full-depth
    relevance (every pool fused, arg-max bytes written by the epilogue)."""
    from cxai.utils.constants import lrp_name_map_6s, LRP_NAME_MAP_TOY
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.explain import lrp_engine
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.xai.drsa.preprocessing import get_intermediate
    if case == "genre_bn":
        net = lrp_ref.genre_model(seed=0, last=256, input_size=(128, 256))
        x = lrp_ref.synth_logmel(3, 128, 256, 20264).cuda()
        comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
        layer, n_fused = net.features[33], 4
    else:
        net = lrp_ref.toy_model(seed=0, last=64)
        x = lrp_ref.synth_logmel(5, 64, 64, 20265).cuda()         # odd batch: the 8 x 8 layer packs two images per tile
        comp = NameMapComposite(LRP_NAME_MAP_TOY)
        layer, n_fused = net.features[13], 3
    plan = lrp_engine._plan(net, comp, x.device)
    assert plan._tc_stack_ok(x)
    res = {}
    for fuse in (True, False):
        plan.fuse_pool = fuse
        res[fuse] = (get_intermediate(net, x, comp, layer, 1), compute_relevances(net, x, comp, class_idx=1))
    plan.fuse_pool = True
    split = plan.module_to_op[layer].index
    H, W = x.shape[2:]
    fused = []
    for k, op in enumerate(plan.ops):
        if op.kind == "conv":
            fused.append(plan._fusable_pool(k, split + 1, split, x.size(0), H, W) is not None)
        elif op.kind == "pool":
            H, W = H // op.kh, W // op.kw
    assert sum(fused) == n_fused
    (a1, R1), full1 = res[True]
    (a0, R0), full0 = res[False]
    assert _rel_per_sample(a1, a0) < 1e-6
    assert _rel_per_sample(R1, R0) < 1e-5
    assert _rel_per_sample(full1, full0) < 1e-5


@pytest.mark.parametrize("B,H,W,Cout", [(3, 13, 37, 64), (2, 8, 32, 40), (1, 1, 5, 8), (2, 64, 64, 64)])
def test_first_layer_ones_backward_matches_generic_path(B, H, W, Cout):
    """WSquare / Flat on the first layer in one pass over NHWC relevance (lrp_tc_first_ones_backward) against the generic
    fp32 pair of convolutions (lrp_conv3x3_backward with x_is_ones) and against fp64 torch; ragged tiles, borders, Cout < 64."""
    L = _L()
    lib = L.lib()
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + W)
    w = torch.randn(Cout, 1, 3, 3, generator=g) ** 2 + 0.01            # WSquare: w^2
    b = torch.randn(Cout, generator=g) ** 2
    R = torch.randn(B, Cout, H, W, generator=g) * (torch.rand(B, Cout, H, W, generator=g) < 0.7)
    eps = 1e-7
    z = F.conv2d(torch.ones(B, 1, H, W, dtype=torch.float64), w.double(), b.double(), padding=1)
    s_ref = R.double() / (z + torch.where(z >= 0, eps, -eps))
    want = F.conv_transpose2d(s_ref, w.double(), padding=1)
    Rn = torch.zeros(B, H, W, 64)
    Rn[..., :Cout] = R.permute(0, 2, 3, 1)
    Rd, wd, bd = Rn.cuda().contiguous(), w.reshape(Cout, 9).cuda().contiguous(), b.cuda()
    out = torch.empty(B, 1, H, W, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.lrp_tc_first_ones_backward(Rd.data_ptr(), wd.data_ptr(), bd.data_ptr(), B, H, W, Cout, 64, eps, out.data_ptr(), st))
    assert _rel_per_sample(out, want) < 1e-5
    # the generic path
    wt = torch.empty(1, Cout, 3, 3, device="cuda")
    w4 = w.cuda().contiguous()
    L.check(lib.lrp_conv3x3_flip_weights(w4.data_ptr(), Cout, 1, wt.data_ptr(), st))
    sbuf = torch.empty(B, Cout, H, W, device="cuda")
    out2 = torch.empty(B, 1, H, W, device="cuda")
    Rc = R.cuda().contiguous()
    L.check(lib.lrp_conv3x3_backward(None, w4.data_ptr(), wt.data_ptr(), bd.data_ptr(), Rc.data_ptr(), B, 1, Cout, H, W, eps, 1,
                                     sbuf.data_ptr(), out2.data_ptr(), st))
    assert _rel_per_sample(out, out2) < 1e-5
    assert lib.lrp_tc_first_ones_backward(Rd.data_ptr(), wd.data_ptr(), bd.data_ptr(), B, H, W, Cout, 128, eps, out.data_ptr(), st) == -2


def test_subspace_filter_kernels_match_fp64():
    """lrp_subspace_project / lrp_subspace_filter (Epsilon on both projections + SubspaceHook mask) vs fp64 torch on
    the SAME inputs, padded leading dimension included."""
    L = _L(); lib = L.lib()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(5)
    for (P, d, m, K, ld) in ((300, 16, 16, 4, 64), (1000, 128, 128, 4, 128), (257, 100, 100, 5, 128), (64, 64, 32, 2, 64)):
        a = torch.relu(torch.randn(P, d, generator=g))
        U = drsa_ref.synth_U0(d, m, seed=d + m)
        ap = torch.zeros(P, ld); ap[:, :d] = a
        ad, Ud = ap.cuda(), U.cuda().contiguous()
        h = torch.empty(P, m, device="cuda"); arec = torch.full((P, ld), float("nan"), device="cuda")
        L.check(lib.lrp_subspace_project(ad.data_ptr(), Ud.data_ptr(), P, d, m, ld, h.data_ptr(), arec.data_ptr(), s))
        # the relevance that reaches InvProjection has the form a' * c (the rule of the layer above multiplies by its input)
        R = torch.randn(P, d, generator=g) * arec.cpu()[:, :d]
        Rp = torch.zeros(P, ld); Rp[:, :d] = R
        Rd = Rp.cuda()
        hw = a.double() @ U.double()
        np.testing.assert_allclose(h.cpu().numpy(), hw.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(arec.cpu()[:, :d].numpy(), (hw @ U.double().T).numpy(), rtol=1e-5, atol=1e-6)
        assert float(arec[:, d:].abs().max()) == 0.0 if ld > d else True
        out = torch.full((K + 1, P, ld), float("nan"), device="cuda")
        ws = torch.empty(int(L.check(lib.lrp_subspace_filter_workspace_bytes(P, d, m))), dtype=torch.uint8, device="cuda")
        L.check(lib.lrp_subspace_filter(ad.data_ptr(), h.data_ptr(), arec.data_ptr(), Rd.data_ptr(), Ud.data_ptr(), P, d, m, K, ld,
                                        1e-6, 1e-6, out.data_ptr(), ws.data_ptr(), ws.numel(), s))
        # fp64 on the kernel's own h and a_rec (the quotients amplify any difference in them, see the test below)
        h64, ar64, U64 = h.cpu().double(), arec.cpu()[:, :d].double(), U.double()
        sref = R.double() / lrp_ref.stabilize(ar64, 1e-6)
        v = h64 * (sref @ U64) / lrp_ref.stabilize(h64, 1e-6)
        d_k = m // K
        got = out.cpu().double()
        tot = torch.zeros(P, d, dtype=torch.float64)
        for k in range(1, K + 1):
            wk = a.double() * (v[:, (k - 1) * d_k:k * d_k] @ U64[:, (k - 1) * d_k:k * d_k].T)
            tot += wk
            assert float((got[k, :, :d] - wk).norm() / wk.norm()) < 2e-5, (P, d, k)
        assert float((got[0, :, :d] - tot).norm() / tot.norm()) < 2e-5
        assert float(got[:, :, d:].abs().max()) == 0.0 if ld > d else True


@pytest.mark.parametrize("case", ["toy", "genre_bn"])
@pytest.mark.parametrize("umode", ["signed_permutation", "orthogonal"])
def test_heatmap_generator_matches_oracle(case, umode):
    """Concept-conditional heatmaps (explainer.py:68-123): HeatmapGenerator on the CUDA engine vs the oracle pushing the
    K+1 clones of every sample through the ProjectionModel like the reference does.

    With a signed permutation U the projections are exact and every map must agree to the LRP tolerance.  With a
    general orthogonal U the concept maps are only defined up to ~1e-2 -- in the reference as well: where a ReLU
    output is exactly 0, a' = (a U) U^T is rounding noise (~1e-8), the relevance arriving there is a' * c, and the
    Epsilon quotient a' c / (a' +- 1e-6) lets ~1 % of c through with the sign of the noise.  The standard map
    (U U^T = I removes it again) and the sum of the concept maps are not affected and are held to the tolerance."""
    from cxai.utils.constants import LRP_NAME_MAP_TOY, lrp_name_map_6s
    from cxai.xai.explain.explainer import HeatmapGenerator, get_class_composite
    from cxai.xai.explain.rules import SequentialMergeBatchNorm
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.model.modify_model import ProjectionModel
    K = 4
    if case == "toy":
        net = lrp_ref.toy_model(seed=0, last=64)
        x = lrp_ref.synth_logmel(5, 64, 64, 11)
        nm, layer_idx, d, cls, canon, shape = LRP_NAME_MAP_TOY, 10, 16, "class2", (), (64, 64)
    else:
        net = lrp_ref.genre_model(seed=0, last=64, input_size=(32, 64))
        x = lrp_ref.synth_logmel(6, 32, 64, 20262)      # the inputs of test_compute_relevances_...: no max-pool near-ties
        nm, layer_idx, d, cls, canon, shape = lrp_name_map_6s(), 26, 128, "blues", [SequentialMergeBatchNorm()], (32, 64)
    if umode == "orthogonal":
        U = drsa_ref.synth_U0(d, seed=40)
    else:
        gperm = torch.Generator().manual_seed(41)
        U = torch.zeros(d, d)
        U[torch.arange(d), torch.randperm(d, generator=gperm)] = torch.where(torch.rand(d, generator=gperm) < 0.5, -1.0, 1.0)
    strict = umode == "signed_permutation"
    gen = HeatmapGenerator(net, U, nm, cls, num_concepts=K, layer_idx=layer_idx, device="cuda", canonizers=canon)
    gen.generate_subspace_heatmaps(x)
    pm = ProjectionModel(net, layer_idx, U.double(), K, case="toy" if case == "toy" else "gtzan")
    comp = get_class_composite(nm, K)
    want = lrp_ref.lrp_pass(pm, x.repeat_interleave(K + 1, dim=0), comp.name_map, lrp_ref.output_modifier(gen.class_idx))
    Hw = want["R_input"].view(x.size(0), K + 1, *shape)
    std = torch.from_numpy(gen.info["standard_heatmaps"])
    assert std.shape == (x.size(0), 1, *shape)
    assert _rel_per_sample(std[:, 0], Hw[:, 0]) < TOL
    # concept heatmaps arrive sorted by descending relevance; undo with the returned mask
    sub = torch.from_numpy(gen.info["subspace_heatmaps"].copy())
    mask = torch.from_numpy(gen.info["mask"].copy())
    rel_w = Hw[:, 1:].sum(dim=(-2, -1))
    if strict:
        np.testing.assert_array_equal(mask.numpy(), torch.argsort(rel_w, dim=-1, descending=True).numpy())
    for b in range(x.size(0)):
        assert sorted(mask[b].tolist()) == list(range(K))
        for j in range(K):
            w = Hw[b, 1 + int(mask[b, j])]
            assert float((sub[b, j].double() - w).norm() / w.norm()) < (TOL if strict else 0.1)
    np.testing.assert_allclose(gen.info["subspace_relevances"], np.take_along_axis(rel_w.numpy(), mask.numpy(), 1),
                               rtol=2e-3 if strict else 0.1, atol=(1e-4 if strict else 2e-2) * float(rel_w.abs().max()))
    assert np.all(np.diff(gen.info["subspace_relevances"], axis=1) <= 0)          # sorted, descending
    # concept maps add up to the standard map (size-independent property)
    assert float((sub.sum(1) - std[:, 0]).abs().max() / std.abs().max()) < 1e-4
    assert gen.info["standard_relevance"].shape == (x.size(0),)
    if not strict:
        return
    # clones that differ (not the reference's use, but defined by the hook): every row keeps the slot of its position
    xr = x.repeat_interleave(K + 1, dim=0).clone()
    xr[1::K + 1] += 0.25
    got = compute_relevances(gen.projectionmodel, xr, gen.composite, class_idx=gen.class_idx)
    wantg = lrp_ref.lrp_pass(pm, xr, comp.name_map, lrp_ref.output_modifier(gen.class_idx))["R_input"]
    ref_scale = wantg.double().flatten(1).norm(dim=1).view(-1, K + 1).max(dim=1).values.repeat_interleave(K + 1)
    err = (got.double().cpu() - wantg).flatten(1).norm(dim=1) / ref_scale
    assert float(err.max()) < TOL


def test_end_to_end_cfg5_small(tmp_path):
    """BASELINE cfg 5 at reduced size: for each class, synthetic log-mel batch -> BatchNorm CNN forward -> LRP to the
    last conv -> (a, c) pairs at all positions -> normalise -> DRSA, against the same pipeline on the CPU oracle;
    also the reference's on-disk formats of the data sets and of the optimiser outputs."""
    import pickle
    from scipy.stats import ortho_group
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.drsa.cluster import optsubspaces, getdrsadata
    net = lrp_ref.genre_model(seed=0, last=64, input_size=(32, 64))
    nm = lrp_name_map_6s()
    comp = NameMapComposite(nm, canonizers=[SequentialMergeBatchNorm()])
    K, steps, layer_idx = 4, 25, 33
    data = {c: lrp_ref.synth_logmel(10, 32, 64, 700 + c) for c in (2, 7)}
    res = optsubspaces.all_classes_pipeline(net, data, comp, layer_idx, str(tmp_path), num_concepts=K, steps=steps, runs=2,
                                            seed=42)
    for c, x in data.items():
        aw, Rw = lrp_ref.get_intermediate(net, x, nm, net.features[layer_idx], c)
        av = drsa_ref.vectors_from_maps_all(aw.float()); rv = drsa_ref.vectors_from_maps_all(Rw.float())
        A = drsa_ref.normalize_vectors(av); C = drsa_ref.normalize_vectors(drsa_ref.compute_context_vectors(av, rv))
        np.random.seed(42)
        U = ortho_group.rvs(A.size(1))
        for run in (1, 2):
            U = U[:, np.random.permutation(A.size(1))]
            objs_ref, U_ref = drsa_ref.run_autograd(A, C, torch.tensor(U, dtype=torch.float32), K, steps)
            root = tmp_path / f"class{c}" / f"layer{layer_idx}" / f"run{run}"
            with open(root / "projection_matrix.pkl", "rb") as f:
                Ug = pickle.load(f)
            lines = open(root / "train_stats.csv").read().splitlines()
            objs = np.array([float(l.split(",")[1]) for l in lines[1:]])
            assert lines[0] == ",loss" and len(objs) == steps + 1
            assert np.max(np.abs(objs - objs_ref) / np.abs(objs_ref)) < 1e-4
            assert drsa_ref.principal_angle(torch.from_numpy(Ug), U_ref, K) < 1e-3
        Ul, hist, rows = res[c]
        assert rows == 10 * 4 and np.allclose(hist, objs)        # 2 x 2 positions at layer 33 of a 32 x 64 input
    # data-set files of getdrsadata.py:26-59 round-trip through the reference's format
    act, ctx = getdrsadata.preprocess_data(net, data[2], comp, layer_idx, class_idx=2, num_locations=None)
    fp = getdrsadata.save_data(act.cpu().numpy(), ctx.cpu().numpy(), layer=layer_idx, sample_class="disco", model="t",
                               output_path=str(tmp_path))
    assert fp.endswith("gtzan/t/disco/dataset_layer33.pkl")
    with open(fp, "rb") as f:
        ds = pickle.load(f)
    assert isinstance(ds, list) and len(ds) == act.size(0) and ds[0][0].shape == (act.size(1),)
    an, cn = getdrsadata.load_and_normalize_data(fp, device="cuda")
    np.testing.assert_allclose(an.cpu().numpy(), drsa_ref.normalize_vectors(act.cpu()).numpy(), rtol=3e-6, atol=1e-8)
    np.testing.assert_allclose(cn.cpu().numpy(), drsa_ref.normalize_vectors(ctx.cpu()).numpy(), rtol=3e-6, atol=1e-7)


def test_prototypes_and_best_run(tmp_path):
    """get_prototypes (prototypes.py:59-130 on an in-memory batch) picks the subset with the highest objective; the
    result-tree reader get_best_run (evaluation.py:107-141) finds the run with the highest final objective."""
    from cxai.utils.constants import LRP_NAME_MAP_TOY
    from cxai.xai.explain.rules import NameMapComposite
    from cxai.xai.drsa.prototypes import get_prototypes
    from cxai.xai.drsa import drsa
    from cxai.utils.evaluation import get_best_run
    net = lrp_ref.toy_model(seed=0, last=64)
    x = lrp_ref.synth_logmel(12, 64, 64, 77)
    U = drsa_ref.synth_U0(64, seed=3)
    K, n = 4, 3
    a, c, idx, objs, _ = get_prototypes(net, 13, U, NameMapComposite(LRP_NAME_MAP_TOY), x, 1, num_concepts=K, n=n, seed=5)
    perm = torch.randperm(12, generator=torch.Generator().manual_seed(5))
    aw, Rw = lrp_ref.get_intermediate(net, x[perm], LRP_NAME_MAP_TOY, net.features[13], 1)
    av = drsa_ref.vectors_from_maps_all(aw.float()); cv = drsa_ref.compute_context_vectors(av, drsa_ref.vectors_from_maps_all(Rw.float()))
    want = [float(drsa_ref.obj_val(av[i * n * 16:(i + 1) * n * 16], cv[i * n * 16:(i + 1) * n * 16], U, K, 16)) for i in range(4)]
    np.testing.assert_allclose(objs, want, rtol=2e-4)
    best = int(np.argmax(want))
    assert idx.tolist() == perm[best * n:(best + 1) * n].tolist()
    assert a.shape == (n * 16, 64) and _rel_per_sample(a, av[best * n * 16:(best + 1) * n * 16]) < 1e-5
    # result tree of drsa.main -> get_best_run
    A, C = drsa_ref.synth_pairs(600, 32, 9)
    drsa.main(A, C, str(tmp_path), num_concepts=2, steps=5, runs=3, seed=1)
    run, loss, _, path, losses = get_best_run(str(tmp_path))
    finals = {r: float(open(tmp_path / f"run{r}" / "train_stats.csv").read().splitlines()[-1].split(",")[1]) for r in (1, 2, 3)}
    assert run == max(finals, key=finals.get) and abs(loss - finals[run]) < 1e-12 and path.endswith(f"run{run}") and len(losses) == 6


def test_graph_replay_of_engine_passes_is_bit_identical(monkeypatch):
    """From the third call with the same (shape, split layer, seed) an engine pass is one CUDA-graph replay
    (LRPPlan.replay_pass); inputs are copied into the static buffer, so new data must give new -- and the same -- results."""
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.explain import lrp_engine
    from cxai.xai.drsa.preprocessing import get_intermediate
    net = lrp_ref.genre_model(seed=0, last=64, input_size=(32, 64))
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    xs = [lrp_ref.synth_logmel(40, 32, 64, 300 + i).cuda() for i in range(4)]
    monkeypatch.setattr(lrp_engine, "USE_GRAPH", False)
    want = [get_intermediate(net, x, comp, net.features[26], 2) for x in xs]
    monkeypatch.setattr(lrp_engine, "USE_GRAPH", True)
    got = [get_intermediate(net, x, comp, net.features[26], 2) for x in xs]        # eager, capture + replay, replay, replay
    plan = lrp_engine._plan(net, comp, xs[0].device)
    assert any(v != "seen" for v in plan._graphs.values())                          # a graph was captured
    for (a, r), (aw, rw) in zip(got, want):
        assert torch.equal(a, aw) and torch.equal(r, rw)
    # full-depth relevance maps (compute_relevances) replay as well
    from cxai.xai.explain.attribute import compute_relevances
    monkeypatch.setattr(lrp_engine, "USE_GRAPH", False)
    Rw = [compute_relevances(net, x, comp, class_idx=c) for x, c in zip(xs, (1, 1, 4, 7))]
    monkeypatch.setattr(lrp_engine, "USE_GRAPH", True)
    Rg = [compute_relevances(net, x, comp, class_idx=c) for x, c in zip(xs, (1, 1, 4, 7))]
    for a, b in zip(Rg, Rw):
        assert torch.equal(a, b)
    # another seed (class) is the same graph with another mask, not a stale replay
    a5, r5 = get_intermediate(net, xs[0], comp, net.features[26], 5)
    monkeypatch.setattr(lrp_engine, "USE_GRAPH", False)
    a5w, r5w = get_intermediate(net, xs[0], comp, net.features[26], 5)
    assert torch.equal(r5, r5w) and not torch.equal(r5, got[0][1])


@pytest.mark.parametrize("num_locations", [None, 5])
def test_context_pairs_from_nhwc_match_the_map_route(num_locations):
    """extract_context_pairs writes the DRSA rows straight from the tensor-core stack's NHWC planes; the result must equal
    get_intermediate -> sample_spatial_locations -> gather_context_pairs (same RNG call order, same rows), also when the
    engine pass is replayed as a CUDA graph."""
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.drsa import preprocessing as pp
    net = lrp_ref.genre_model(seed=0, last=64, input_size=(32, 64))
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    for rep in range(3):                                      # plain launches, capture, replay
        x = lrp_ref.synth_logmel(40, 32, 64, 400 + rep).cuda()
        np.random.seed(9)
        a_maps, R_maps = pp.get_intermediate(net, x, comp, net.features[26], 2)
        idcs = pp.sample_spatial_locations(40, tuple(a_maps.shape[-2:]), num_locations) if num_locations else None
        act_w, ctx_w = pp.gather_context_pairs(a_maps, R_maps, idcs, normalize=True)
        np.random.seed(9)
        act, ctx = pp.extract_context_pairs(net, x, comp, 26, 2, num_locations=num_locations, normalize=True)
        assert act.shape == act_w.shape == (40 * (num_locations or a_maps.shape[-1] * a_maps.shape[-2]), 128)
        np.testing.assert_allclose(act.cpu().numpy(), act_w.cpu().numpy(), rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(ctx.cpu().numpy(), ctx_w.cpu().numpy(), rtol=2e-6, atol=1e-9)
    # a split layer below the tensor-core stack's reach (or the toy model on the fp32 kernels) takes the map route
    toy = lrp_ref.toy_model(seed=0, last=64)
    from cxai.utils.constants import LRP_NAME_MAP_TOY
    xt = lrp_ref.synth_logmel(6, 64, 64, 3).cuda()
    act, ctx = pp.extract_context_pairs(toy, xt, NameMapComposite(LRP_NAME_MAP_TOY), 13, 0)
    a_maps, R_maps = pp.get_intermediate(toy, xt, NameMapComposite(LRP_NAME_MAP_TOY), toy.features[13], 0)
    act_w, ctx_w = pp.gather_context_pairs(a_maps, R_maps, None, normalize=True)
    np.testing.assert_allclose(act.cpu().numpy(), act_w.cpu().numpy(), rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(ctx.cpu().numpy(), ctx_w.cpu().numpy(), rtol=2e-6, atol=1e-9)
