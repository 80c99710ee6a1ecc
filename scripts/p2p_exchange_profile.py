"""Cost of the fused peer exchange: %globaltimer stamps of CTA 0 of the finish kernel with and without it.
   torchrun --nproc-per-node N scripts/p2p_exchange_profile.py      (N = 2 .. 8)"""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drsa_audio_b200 import _lib as L
from cxai.xai.drsa.drsa import SubspaceOptimizer
from bench import synth_rows_cuda
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d, M, K = 256, 640000, 4
A, C = synth_rows_cuda(M, d, 1 + rank, dev)
U0 = torch.linalg.qr(torch.randn(d, d, generator=torch.Generator().manual_seed(3)))[0]
for exch in ("p2p", "nccl"):
    opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, precision="tc", use_cuda_graph=False, exchange=exch)
    opt._rows.split_u(opt._Uw)
    opt.reset_log(256)
    for _ in range(5): opt._step(opt._obj_log, -1, True)
    buf = torch.zeros(64, dtype=torch.int64, device=dev)
    L.lib().drsa_debug_set_tc_profile(buf.data_ptr())
    first, total, ev = [], [], []
    for it in range(20):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        opt._rows.step(opt._Uw)
        e1.record()
        if opt._px is None:
            dist.all_reduce(opt._rows.sums)
        opt._rows.finish(opt._Uw, opt.M_global, opt._obj_log, -1, True, opt.retraction_iters, opt.retraction_tol, opt._px)
        e2.record()
        torch.cuda.synchronize()
        v = buf.cpu().tolist(); n = v[15]; st = v[16:16 + n]
        first.append((st[1] - st[0]) / 1e3); total.append((st[-1] - st[0]) / 1e3); ev.append((e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
        deltas = [(b - a) / 1e3 for a, b in zip(st[:-1], st[1:])]
    L.lib().drsa_debug_set_tc_profile(None)
    med = lambda x: sorted(x)[len(x) // 2]
    print(f"rank {rank} {exch}: phase 0 (incl. exchange) median {med(first):.1f} us, finish kernel {med(total):.1f} us, "
          f"row pass {med([a for a, _ in ev]):.1f} us, after row pass (events) {med([b for _, b in ev]):.1f} us", flush=True)
    if rank == 0:
        print(f"rank 0 {exch}: phases of the last finish kernel (us): " + " ".join(f"{v:.1f}" for v in deltas), flush=True)
    del opt
dist.destroy_process_group()
