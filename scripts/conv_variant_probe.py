"""Accuracy and speed of the tensor-core conv stack when the x_lo * w_hi product is skipped (lrp_debug_set_conv_variant) in
the forward / ratio / input-multiply launches, against the reference fixtures (full-resolution cfg-2 CNN, arch A small)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from drsa_audio_b200 import _lib as L
from cxai.model.create_model import VGGType
from cxai.utils.constants import lrp_name_map_6s
from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
from cxai.xai.explain import lrp_engine
from cxai.xai.explain.attribute import compute_relevances
from cxai.xai.drsa import preprocessing as pp

def rel(got, want):
    got, want = torch.as_tensor(got).double().cpu().flatten(1), torch.as_tensor(want).double().flatten(1)
    return float(((got - want).norm(dim=1) / want.norm(dim=1)).max())

gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
lrp_engine.USE_GRAPH = False
for variant in (0, 1, 2, 4, 3, 7):
    L.lib().lrp_debug_set_conv_variant(variant)
    out = []
    for name, H, W in (("cfg2_full", 128, 256), ("archA_small", 32, 64)):
        g = np.load(os.path.join(gold, f"lrp_{name}.npz"))
        net = synth.build_model(VGGType, str(g["model"]), 0, 1)
        x = synth.synth_logmel(int(g["N"]), H, W, int(g["x_seed"])).cuda()
        comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
        a, R = pp.get_intermediate(net, x, comp, net.features[33], 3)
        Rin = compute_relevances(net, x, comp, class_idx=3)
        errs = [rel(a, g["a_l33_f64"]), rel(R, g["R_l33_f64"])]
        rin = ((Rin.double().cpu().flatten(1) - torch.as_tensor(g["Rin_c3_f64"]).double().flatten(1)).norm(dim=1) /
               torch.as_tensor(g["Rin_c3_f64"]).double().flatten(1).norm(dim=1)).numpy()
        out.append(f"{name}: a {errs[0]:.1e} R {errs[1]:.1e} Rin {np.array2string(rin, precision=1)}")
        if name == "archA_small":
            for layer in (19, 26):
                a2, R2 = pp.get_intermediate(net, x, comp, net.features[layer], 3)
                out.append(f"l{layer}: a {rel(a2, g[f'a_l{layer}_f64']):.1e} R {rel(R2, g[f'R_l{layer}_f64']):.1e}")
    # speed: 256 samples through extract_context_pairs on the cfg-2 CNN
    net = synth.build_model(VGGType, "cfg2", 0, 1).cuda()
    xb = synth.synth_logmel(256, 128, 256, 5).cuda()
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    for _ in range(2):
        pp.extract_context_pairs(net, xb, comp, 33, 0)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pp.extract_context_pairs(net, xb, comp, 33, 0); e1.record(); torch.cuda.synchronize()
    print(f"variant {variant}: {e0.elapsed_time(e1):.3f} ms per 256 samples | " + " | ".join(out), flush=True)
L.lib().lrp_debug_set_conv_variant(0)
