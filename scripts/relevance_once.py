"""Full-depth LRP relevance maps (compute_relevances) at the cfg-2 CNN for N samples; --ncu: one warm + one measured pass."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_cfg2_model
from cxai.utils.constants import lrp_name_map_6s
from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
from cxai.xai.explain.attribute import compute_relevances
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 128
net = build_cfg2_model(dev)
comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
g = torch.Generator(device=dev).manual_seed(20262)
x = (1.2 * torch.randn(n, 1, 128, 256, generator=g, device=dev) - 1.5).clamp(min=-4.0)
compute_relevances(net, x, comp, class_idx=0); torch.cuda.synchronize()
if "--ncu" in sys.argv:
    compute_relevances(net, x, comp, class_idx=0); torch.cuda.synchronize(); sys.exit(0)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); compute_relevances(net, x, comp, class_idx=0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"{n} samples: {sorted(ts)[2]:.3f} ms -> {n / sorted(ts)[2] * 1e3:.0f} maps/s")
