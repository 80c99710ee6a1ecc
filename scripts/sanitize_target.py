"""Small shapes through every hand-rolled mbarrier / TMEM pipeline, for compute-sanitizer (scripts/sanitize.sh):
the DRSA row pass in every arithmetic mode (d = 128, 256, 512), the fused finish kernel, and the tensor-core convolution
stack (forward with fused max-pool, ratio / input-multiply backward epilogues) on a reduced BatchNorm model."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import drsa_ref, synth
from cxai.xai.drsa.drsa import SubspaceOptimizer

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "drsa"):
    for d, K in ((128, 4), (256, 4), (512, 8)):
        A, C = drsa_ref.synth_pairs(1000, d, 3)          # ragged: 1000 rows = 15 subtiles + a partial one
        U0 = drsa_ref.synth_U0(d, d, 4)
        for prec in ("tc", "tc_split", "tc_hilo", "tc_dc", "tc32", "fp32"):
            if d == 512 and prec in ("tc_split", "tc32"):
                continue
            opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device="cuda", precision=prec, use_cuda_graph=False)
            opt.run(steps=3, save=False)
            torch.cuda.synchronize()
            print("drsa", d, prec, float(opt.obj_history[-1]), flush=True)
if what in ("all", "lrp"):
    from cxai.model.create_model import VGGType
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.explain.attribute import compute_relevances
    from cxai.xai.drsa.preprocessing import get_intermediate
    net = synth.build_model(VGGType, "archA_small", 0, 1)
    x = synth.synth_logmel(3, 32, 64, 5).cuda()
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    a, R = get_intermediate(net, x, comp, net.features[26], 3)
    Rin = compute_relevances(net, x, comp, class_idx=3)
    torch.cuda.synchronize()
    print("lrp", float(a.sum()), float(R.sum()), float(Rin.sum()), flush=True)
