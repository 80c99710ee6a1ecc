"""2-GPU NCCL test of the row-sharded DRSA optimiser (skipped with fewer than 2 GPUs): two ranks each hold
half of the rows; the trajectory must match the single-GPU run on all rows and the replicas of U must be
bit-identical."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import drsa_ref
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    M, d, K, steps = 20000, 128, 4, 12
    A, C = drsa_ref.synth_pairs(M, d, 77)
    U0 = drsa_ref.synth_U0(d, seed=78)
    lo, hi = (0, 9000) if rank == 0 else (9000, M)           # uneven shards
    out = {}
    # exchange: NCCL all-reduce between the kernels / fused into the finish kernel over peer memory (CUDA graph replay
    # and direct launches)
    for prec, exch, graph in (("fp32", "nccl", True), ("tc", "nccl", True), ("fp32", "p2p", True), ("tc", "p2p", True),
                              ("tc", "p2p", False), ("tc_dc", "p2p", True), ("tc_hilo", "p2p", True), ("tc_dc", "nccl", False)):
        opt = SubspaceOptimizer(U0, A[lo:hi], C[lo:hi], None, num_concepts=K, device=f"cuda:{rank}", precision=prec,
                                exchange=exch, use_cuda_graph=graph)
        assert opt.exchange == exch and opt.use_cuda_graph == (graph and exch == "p2p")
        opt.run(steps=steps, save=False)
        gathered = [torch.zeros_like(opt.U) for _ in range(world)]
        dist.all_gather(gathered, opt.U)
        out[(prec, exch, graph)] = (opt.obj_history.copy(), opt.U.cpu(), float((gathered[0] - gathered[1]).abs().max()),
                                    opt.M_global)
    # small split layers (d <= 64: cfg 1, the toy CNN) take the single-CTA finish kernel, which carries the exchange as well
    As, Cs = drsa_ref.synth_pairs(6000, 64, 79)
    Us = drsa_ref.synth_U0(64, seed=80)
    small = {}
    for exch in ("p2p", "nccl"):
        o = SubspaceOptimizer(Us, As[3500 * rank:3500 * (rank + 1)], Cs[3500 * rank:3500 * (rank + 1)], None, num_concepts=K,
                              device=f"cuda:{rank}", precision="fp32", exchange=exch)
        o.run(steps=steps, save=False)
        g2 = [torch.zeros_like(o.U) for _ in range(world)]
        dist.all_gather(g2, o.U)
        small[exch] = (o.obj_history.copy(), o.U.cpu(), float((g2[0] - g2[1]).abs().max()))
    if rank == 0:
        objs_s, U_s = drsa_ref.run_autograd(As, Cs, Us, K, steps)
        for exch, (objs, U, rep) in small.items():
            assert rep == 0.0 and float(np.max(np.abs(objs - objs_s) / np.abs(objs_s))) < 1e-5, (exch, rep)
            assert drsa_ref.principal_angle(U, U_s, K) < 1e-4
        objs_ref, U_ref = drsa_ref.run_autograd(A, C, U0, K, steps)
        for key, (objs, U, rep, Mg) in out.items():
            ret["/".join(map(str, key))] = (float(np.max(np.abs(objs - objs_ref) / np.abs(objs_ref))),
                                            drsa_ref.principal_angle(U, U_ref, K), rep, Mg)
        # two ranks: a + b is the same float whoever adds it, so both exchanges give the same bits
        ret["same_bits"] = bool(torch.equal(out[("tc", "nccl", True)][1], out[("tc", "p2p", True)][1])
                                and torch.equal(out[("tc", "p2p", True)][1], out[("tc", "p2p", False)][1])
                                and torch.equal(out[("fp32", "nccl", True)][1], out[("fp32", "p2p", True)][1]))
    dist.destroy_process_group()


def test_two_rank_row_sharding_matches_reference():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 1000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert len(ret) == 9
        for key in ret.keys():
            if key == "same_bits":
                continue
            rel, ang, rep, Mg = ret[key]
            print(key, rel, ang, rep)
            assert Mg == 20000
            assert rel < 1e-4 and ang < 1e-3
            assert rep == 0.0
        assert ret["same_bits"]


def _pipeline_worker(rank, world, port, ret):
    """cfg-5 pipeline with the classes spread over the ranks (plan_class_schedule) against the row-sharded schedule."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import drsa_ref, synth
    from cxai.model.create_model import VGGType
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.drsa.cluster.optsubspaces import all_classes_pipeline
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    net = synth.build_model(VGGType, "archA_small", 0, 1).to(dev)
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    classes = 3                                              # one full round (2 classes) + a partial round (1 class on 2 ranks)
    n = (24, 40)[rank]                                       # uneven sample counts per rank
    data = {c: synth.synth_logmel(n, 32, 64, 500 + 10 * c + rank).to(dev) for c in range(classes)}
    res = {}
    for schedule in ("classes", "shard"):
        out = all_classes_pipeline(net, data, comp, 26, None, num_concepts=4, steps=30, schedule=schedule, device=dev,
                                   precision="fp32")
        res[schedule] = {c: (out[c][0].cpu(), np.asarray(out[c][1])) for c in range(classes)}
    gathered = [None] * world
    dist.all_gather_object(gathered, {c: res["classes"][c][0] for c in range(classes)})
    if rank == 0:
        ok_rep = all(torch.equal(gathered[0][c], gathered[1][c]) for c in range(classes))       # every rank ends with every U
        worst_obj = worst_ang = 0.0
        for c in range(classes):
            Ua, oa = res["classes"][c]
            Ub, ob = res["shard"][c]
            worst_obj = max(worst_obj, float(np.max(np.abs(oa - ob) / np.abs(ob))))
            worst_ang = max(worst_ang, drsa_ref.principal_angle(Ua, Ub, 4))
        ret["pipeline"] = (ok_rep, worst_obj, worst_ang, [len(res["classes"][c][1]) for c in range(classes)])
    dist.destroy_process_group()


def test_two_rank_class_schedule_matches_row_sharding():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 1000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_pipeline_worker, args=(2, port, ret), nprocs=2, join=True)
        ok_rep, worst_obj, worst_ang, lens = ret["pipeline"]
        print("class schedule vs row sharding:", worst_obj, worst_ang)
        assert ok_rep and lens == [31, 31, 31]
        assert worst_obj < 1e-5 and worst_ang < 1e-4          # same rows, same arithmetic, another summation order
