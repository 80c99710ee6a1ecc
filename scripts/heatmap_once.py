"""Concept-conditional heatmaps (HeatmapGenerator.generate_subspace_heatmaps, explainer.py:68-123) at the cfg-2 CNN:
N samples, K = 4 concepts at features[33] (d = 256) -> N standard + N x K concept heatmaps of 128 x 256."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_cfg2_model
from cxai.utils.constants import lrp_name_map_6s
from cxai.xai.explain.rules import SequentialMergeBatchNorm
from cxai.xai.explain.explainer import HeatmapGenerator
from bench import synth_U0
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = build_cfg2_model(dev)
U = synth_U0(256, seed=5)
gen = HeatmapGenerator(net, U, lrp_name_map_6s(), "blues", num_concepts=4, layer_idx=33, device=dev,
                       canonizers=[SequentialMergeBatchNorm()])
g = torch.Generator(device=dev).manual_seed(20262)
x = (1.2 * torch.randn(n, 1, 128, 256, generator=g, device=dev) - 1.5).clamp(min=-4.0)
gen.generate_subspace_heatmaps(x); torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); gen.generate_subspace_heatmaps(x); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t = sorted(ts)[1]
h = gen.info["subspace_heatmaps"]
print(f"{n} samples: {t * 1e3:.1f} ms wall (incl. D2H of {n * 5} maps) -> {n * 5 / t:.0f} heatmaps/s; "
      f"sum of concept maps vs standard map: max rel diff {abs(h.sum(1) - gen.info['standard_heatmaps'][:, 0]).max() / abs(gen.info['standard_heatmaps']).max():.2e}")
