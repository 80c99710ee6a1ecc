"""Stage 1 at the cfg-2 CNN: times get_intermediate + gather for N samples with the conv+pool fusion on and off;
with --ncu runs one warm pass and one measured pass only (for an ncu launch list)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_cfg2_model
from cxai.utils.constants import lrp_name_map_6s
from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
from cxai.xai.explain import lrp_engine
from cxai.xai.drsa import preprocessing as pp
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
net = build_cfg2_model(dev)
comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
g = torch.Generator(device=dev).manual_seed(20262)
x = (1.2 * torch.randn(n, 1, 128, 256, generator=g, device=dev) - 1.5).clamp(min=-4.0)
layer = net.features[33]
def once():
    a, R = pp.get_intermediate(net, x, comp, layer, 0)
    return pp.gather_context_pairs(a, R, None, normalize=True)
lrp_engine.ENGINE_CHUNK = int(os.environ.get('CHUNK', lrp_engine.ENGINE_CHUNK))
plan = lrp_engine._plan(net, comp, dev)
if "--ncu" in sys.argv:
    once(); torch.cuda.synchronize(); once(); torch.cuda.synchronize(); sys.exit(0)
for fuse in (True, False, True):
    plan.fuse_pool = fuse
    once(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); once(); e1.record(); t_host = time.perf_counter() - t0
        torch.cuda.synchronize(); ts.append((e0.elapsed_time(e1), t_host * 1e3))
    print(f"fuse_pool={fuse}: {n} samples, device ms {sorted(t for t, _ in ts)[2]:.3f}, host enqueue ms {sorted(h for _, h in ts)[2]:.3f}", flush=True)
