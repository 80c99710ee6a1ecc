"""2-GPU debug aid: after every step compare the all-reduced sums and the replicas of U bit for bit."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import drsa_ref
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    M, d, K, steps = 20000, 128, 4, 12
    A, C = drsa_ref.synth_pairs(M, d, 77)
    U0 = drsa_ref.synth_U0(d, seed=78)
    lo, hi = (0, 9000) if rank == 0 else (9000, M)
    for prec in ("fp32", "tc"):
        opt = SubspaceOptimizer(U0, A[lo:hi], C[lo:hi], None, num_concepts=K, device=f"cuda:{rank}", precision=prec)
        opt._rows.split_u(opt.U)
        opt.reset_log(steps + 1)
        for s in range(steps):
            opt._rows.step(opt.U)
            local = opt._rows.sums.clone()
            dist.all_reduce(opt._rows.sums)
            g = [torch.zeros_like(opt._rows.sums) for _ in range(world)]
            dist.all_gather(g, opt._rows.sums)
            gl = [torch.zeros_like(local) for _ in range(world)]
            dist.all_gather(gl, local)
            manual = gl[0] + gl[1]
            ds = float((g[0] - g[1]).abs().max())
            dm = float((g[rank] - manual).abs().max())
            opt._rows.finish(opt.U, opt.M_global, opt._obj_log, -1, True, opt.retraction_iters, opt.retraction_tol)
            gu = [torch.zeros_like(opt.U) for _ in range(world)]
            dist.all_gather(gu, opt.U)
            du = float((gu[0] - gu[1]).abs().max())
            st = opt._rows.status.cpu().tolist()
            gs = [None] * world
            dist.all_gather_object(gs, st)
            if rank == 0:
                print(f"{prec} step {s}: sums rank diff {ds:.3e}  nccl-vs-manual {dm:.3e}  U rank diff {du:.3e}  status {gs}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    mp.spawn(worker, args=(2, 29711), nprocs=2, join=True)
