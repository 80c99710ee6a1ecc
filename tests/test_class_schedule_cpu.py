"""cfg-5 scheduling on the host (gloo, world size 3): classes are independent problems that are spread over the ranks
(``plan_class_schedule``) and their rows, produced sample-sharded by stage 1, are moved to the ranks that optimise them by
one all-to-all per round (``redistribute_rows``).  Checked: every class is scheduled exactly once, groups of a round are
disjoint, and after the move the ranks of a class hold every row of that class exactly once."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_covers_every_class_once_with_disjoint_groups():
    sys.path.insert(0, ROOT)
    from cxai.xai.drsa.cluster.optsubspaces import plan_class_schedule
    for classes, world in [(10, 8), (10, 1), (10, 4), (3, 8), (8, 8), (16, 8), (5, 3), (1, 2)]:
        plan = plan_class_schedule(classes, world)
        seen = []
        for rnd in plan:
            ranks = [r for _, g in rnd for r in g]
            assert len(ranks) == len(set(ranks)) and all(0 <= r < world for r in ranks)
            seen += [c for c, _ in rnd]
        assert sorted(seen) == list(range(classes))
    assert plan_class_schedule(10, 8) == [[(j, [j]) for j in range(8)], [(8, [0, 1, 2, 3]), (9, [4, 5, 6, 7])]]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from cxai.xai.drsa.cluster.optsubspaces import plan_class_schedule, redistribute_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    classes, d = 5, 4
    # rank p holds (3 + p + c) rows of class c; row value encodes (class, source rank, row index)
    local = {c: torch.tensor([[c, rank, i, 0.0] for i in range(3 + rank + c)], dtype=torch.float32) for c in range(classes)}
    counts = torch.tensor([[3 + p + c for c in range(classes)] for p in range(world)], dtype=torch.int64)
    got = {}
    for rnd in plan_class_schedule(classes, world):
        rows = redistribute_rows(local, rnd, counts, rank, world)
        mine = next((c for c, g in rnd if rank in g), None)
        assert (rows is None) == (mine is None)
        if rows is not None:
            assert bool((rows[:, 0] == mine).all())
            got[mine] = rows
    gathered = [None] * world
    dist.all_gather_object(gathered, {c: v.numpy().tolist() for c, v in got.items()})
    if rank == 0:
        ok = True
        for c in range(classes):
            rows = [tuple(r) for g in gathered for r in g.get(c, [])]
            want = [(float(c), float(p), float(i), 0.0) for p in range(world) for i in range(3 + p + c)]
            ok = ok and sorted(rows) == sorted(want)
        ret["ok"] = ok
    dist.destroy_process_group()


def test_rows_reach_their_class_owners_exactly_once():
    port = 29500 + ((os.getpid() + 7) % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(3, port, ret), nprocs=3, join=True)
        assert ret["ok"]
