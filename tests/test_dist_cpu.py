"""world_size-2 gloo test of the data-parallel protocol of the DRSA step: rows shard across ranks,
U is replicated, ONE all-reduce of d*m+K floats joins the row sums, every rank then derives the same
objective / gradient.  The per-rank row pass is played by the CPU oracle here (the CUDA kernels are
exercised by the gpu-marked tests); what is under test is the sharding + reduction logic."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from oracle import drsa_ref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    M, d, K = 1001, 32, 4                         # ragged split on purpose
    A, C = drsa_ref.synth_pairs(M, d, 31)
    U = drsa_ref.synth_U0(d, seed=32).double()
    bounds = np.linspace(0, M, world + 1).astype(int)
    lo, hi = bounds[rank], bounds[rank + 1]
    for _ in range(3):
        X, ss = drsa_ref.step_sums(A[lo:hi].double(), C[lo:hi].double(), U, K)
        buf = torch.cat([X.reshape(-1), ss])      # the d*m + K floats of the protocol
        dist.all_reduce(buf)
        Mg = torch.tensor([hi - lo]); dist.all_reduce(Mg)
        obj, grad = drsa_ref.finish_from_sums(buf[: d * d].view(d, d), buf[d * d:], int(Mg), K)
        U = drsa_ref.orthogonalize(U + grad)
    gathered = [torch.zeros_like(U) for _ in range(world)]
    dist.all_gather(gathered, U)
    if rank == 0:
        # replicas stay identical and equal the single-process result
        Us = drsa_ref.synth_U0(d, seed=32).double()
        for _ in range(3):
            _, _, Us = drsa_ref.step_closed_form(A.double(), C.double(), Us, K)
        ret["replica_diff"] = float((gathered[0] - gathered[1]).abs().max())
        ret["single_diff"] = float((gathered[0] - Us).abs().max())
        ret["Mg"] = int(Mg)
    dist.destroy_process_group()


def test_row_sharded_step_matches_single_process():
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["Mg"] == 1001
        assert ret["replica_diff"] == 0.0
        assert ret["single_diff"] < 1e-12
