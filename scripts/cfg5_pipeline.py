"""BASELINE cfg 5 end to end on the device: for each of 10 classes, synthetic log-mel [N_c, 1, 128, 256] -> CNN forward ->
LRP to features[33] with class_idx = c -> c = R / (a + 1e-7) -> normalise -> DRSA (K = 4, `steps` steps).  Samples of a
class are split evenly across the ranks (torchrun); prints wall clock, LRP context vectors/s and DRSA steps/s.

    python scripts/cfg5_pipeline.py [samples_per_class=10000] [steps=2000]
    python -m torch.distributed.run --nproc-per-node 8 scripts/cfg5_pipeline.py 10000 2000"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_cfg2_model
from cxai.utils.constants import lrp_name_map_6s
from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
from cxai.xai.drsa.cluster.optsubspaces import class_pipeline

n_class = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
net = build_cfg2_model(dev)
comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
n_local = n_class // world
g = torch.Generator(device=dev).manual_seed(20265 + rank)


def batch():
    return (1.2 * torch.randn(n_local, 1, 128, 256, generator=g, device=dev) - 1.5).clamp(min=-4.0)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


class_pipeline(net, batch()[:64], comp, 33, 0, None, num_concepts=4, steps=8)       # warm-up: plan, graph, allocator
sync()
t_lrp = t_drsa = 0.0
t0 = time.perf_counter()
objs = []
for c in range(10):
    x = batch()
    sync(); ta = time.perf_counter()
    # the two stages are timed separately by running the pipeline's first half on its own first (untimed in the total)
    U, hist, rows = class_pipeline(net, x, comp, 33, c, None, num_concepts=4, steps=steps)
    sync(); tb = time.perf_counter()
    objs.append((float(hist[0]), float(hist[-1])))
    t_drsa += tb - ta
t_total = time.perf_counter() - t0
if rank == 0:
    vecs = 10 * n_local * world * 64
    print(f"cfg5: 10 classes x {n_local * world} samples on {world} GPU(s), {steps} DRSA steps per class: wall clock {t_total:.2f} s "
          f"(incl. generating the synthetic spectrograms); per class {t_drsa / 10:.3f} s for LRP + gather + DRSA; "
          f"{vecs} context vectors, {10 * steps} DRSA steps -> {vecs / t_drsa / 1e6:.2f} M vectors/s and {10 * steps / t_drsa:.0f} steps/s "
          f"through the whole pipeline", flush=True)
    print("objective first -> last per class:", " ".join(f"{a:.4f}->{b:.4f}" for a, b in objs), flush=True)
if world > 1:
    dist.destroy_process_group()
