"""Per-class subspace optimisation drivers -- mirror of the reference's cxai/xai/drsa/cluster/optsubspaces.py
(:8-52: for every class and split layer load the stored pairs, normalise, ``drsa.main``), plus the BASELINE cfg-5
pipeline that never leaves the device: spectrograms -> CNN forward -> LRP to the split layer -> (a, c) pairs ->
normalise -> DRSA, once per class.  With ``torch.distributed`` initialised every rank passes ITS samples of the class:
the LRP pass needs no communication, the normalisation statistics and the d*m + K row sums of every DRSA step are
all-reduced (rows never move), and every rank ends with the same U."""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from cxai.xai.drsa import drsa
from cxai.xai.drsa import preprocessing as pp
from cxai.xai.drsa.cluster.getdrsadata import load_and_normalize_data

__all__ = ["optimize_stored_classes", "class_pipeline", "all_classes_pipeline", "plan_class_schedule", "redistribute_rows"]


def optimize_stored_classes(path_to_data: str, path_to_models: str, class_idx_mapper: Dict[str, int],
                            layer_idcs: Iterable[int], num_concepts: int = 4, steps: int = 5000, runs: int = 3,
                            seed: int = 42, device="cuda", **optimizer_kwargs) -> None:
    """optsubspaces.main (:8-52) with the paths as arguments."""
    for sample_class in class_idx_mapper.keys():
        for layer_idx in layer_idcs:
            a, c = load_and_normalize_data(os.path.join(path_to_data, f"{sample_class}/dataset_layer{layer_idx}.pkl"),
                                           device=device)
            drsa.main(a, c, os.path.join(path_to_models, sample_class, f"layer{layer_idx}"), num_concepts=num_concepts,
                      steps=steps, runs=runs, seed=seed, device=device, **optimizer_kwargs)


def class_pipeline(model, input_batch: torch.Tensor, composite, layer_idx: int, class_idx: int, model_root: Optional[str],
                   num_concepts: int = 4, steps: int = 2000, runs: int = 1, seed: int = 42,
                   num_locations: Optional[int] = None, device="cuda", **optimizer_kwargs):
    """One class of BASELINE cfg 5 on the device.  Returns (U of the last run, objective history of the last run,
    number of rows on this rank)."""
    from scipy.stats import ortho_group
    dev = torch.device(device)
    act, ctx = pp.extract_context_pairs(model, input_batch, composite, layer_idx, class_idx, num_locations=num_locations,
                                        normalize=True, device=dev)
    np.random.seed(seed)                        # same start on every rank (drsa.py:263-272)
    d = act.size(-1)
    U = ortho_group.rvs(d)
    opt = None
    is_rank0 = not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
        torch.distributed.get_rank() == 0
    for run in range(1, runs + 1):
        U = U[:, np.random.permutation(d)]
        path = None
        if model_root is not None and is_rank0:
            path = os.path.join(model_root, f"run{run}")
            os.makedirs(path, exist_ok=True)
        opt = drsa.SubspaceOptimizer(torch.tensor(U, dtype=torch.float32), act, ctx, path, num_concepts=num_concepts,
                                     device=dev, **optimizer_kwargs)
        opt.run(steps=steps, save=path is not None)
    return opt.U, opt.obj_history, act.size(0)


def plan_class_schedule(num_classes: int, world: int) -> List[List[Tuple[int, List[int]]]]:
    """Rounds of (class position, ranks that optimise it).  Classes are INDEPENDENT problems (SURVEY 8e "alternatives"), and
    a DRSA step on a small row shard is latency-bound (80 000 rows: 45 us of row pass under a 60 us finish kernel), so
    classes are spread over the GPUs instead of every class over all GPUs: full rounds give every rank one class of its own
    (no exchange at all); a last, partial round splits the ranks into as many groups as classes are left and shards the rows
    of each inside its group."""
    rounds, pos = [], 0
    for _ in range(num_classes // world):
        rounds.append([(pos + j, [j]) for j in range(world)])
        pos += world
    rem = num_classes - pos
    if rem:
        sizes = [world // rem + (1 if i < world % rem else 0) for i in range(rem)]
        start, last = 0, []
        for i, n in enumerate(sizes):
            last.append((pos + i, list(range(start, start + n))))
            start += n
        rounds.append(last)
    return rounds


def _split_counts(n: int, parts: int) -> List[int]:
    return [n // parts + (1 if i < n % parts else 0) for i in range(parts)]


def redistribute_rows(local: Dict[int, torch.Tensor], round_plan: List[Tuple[int, List[int]]], counts: torch.Tensor,
                      rank: int, world: int, group=None) -> Optional[torch.Tensor]:
    """One all-to-all that moves the rows of the classes of a round from where stage 1 produced them (every rank holds the
    rows of ITS samples of every class) to the ranks that optimise them.  ``local[c]``: this rank's [n, d] rows of class
    position c; ``counts[p, c]``: rows rank p holds of class c (known everywhere).  Every rank of a class's group receives
    an equal share of every source rank's rows.  Returns this rank's rows of its class of the round (None if it idles)."""
    dist = torch.distributed
    dest_class = {}                      # dest rank -> (class position, index in group, group size)
    for c, ranks in round_plan:
        for i, r in enumerate(ranks):
            dest_class[r] = (c, i, len(ranks))
    any_t = next(iter(local.values()))
    d, dev, dt = any_t.size(1), any_t.device, any_t.dtype
    send, in_split = [], []
    for q in range(world):
        if q not in dest_class:
            in_split.append(0)
            continue
        c, i, gs = dest_class[q]
        parts = _split_counts(int(counts[rank, c]), gs)
        r0 = sum(parts[:i])
        send.append(local[c][r0:r0 + parts[i]])
        in_split.append(parts[i])
    out_split = [0] * world
    if rank in dest_class:
        c, i, gs = dest_class[rank]
        out_split = [_split_counts(int(counts[p, c]), gs)[i] for p in range(world)]
    sbuf = torch.cat(send, 0) if send else torch.empty(0, d, device=dev, dtype=dt)
    rbuf = torch.empty(sum(out_split), d, device=dev, dtype=dt)
    dist.all_to_all_single(rbuf, sbuf.contiguous(), output_split_sizes=out_split, input_split_sizes=in_split, group=group)
    return rbuf if rank in dest_class else None


_SUBGROUPS: Dict[Tuple[int, ...], object] = {}


def _subgroup(ranks: List[int]):
    """torch.distributed group of ``ranks`` (created collectively by ALL ranks, cached)."""
    key = tuple(ranks)
    if key not in _SUBGROUPS:
        _SUBGROUPS[key] = torch.distributed.new_group(list(ranks))
    return _SUBGROUPS[key]


def all_classes_pipeline(model, data_by_class: Dict[int, torch.Tensor], composite, layer_idx: int, model_root: Optional[str],
                         num_concepts: int = 4, steps: int = 2000, runs: int = 1, seed: int = 42,
                         num_locations: Optional[int] = None, device="cuda", schedule: str = "auto",
                         timings: Optional[dict] = None, **optimizer_kwargs):
    """BASELINE cfg 5: the per-class pipeline for every class (class index -> this rank's spectrograms of that class).

    Single process, or ``schedule='shard'``: class after class, the rows of a class sharded over all ranks (one exchange
    of d*m + K floats per step).  ``schedule='classes'`` (default under torch.distributed): stage 1 runs sample-sharded
    for every class (no communication apart from the normalisation statistics), then the rows are moved over NVLink to
    the ranks of ``plan_class_schedule`` and whole classes are optimised side by side.  Every rank returns every class's
    (U, objective history, rows optimised on this rank)."""
    dist = torch.distributed
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if schedule == "auto":
        schedule = "classes" if distributed else "shard"
    if not distributed or schedule == "shard":
        out = {}
        for class_idx, batch in data_by_class.items():
            root = None if model_root is None else os.path.join(model_root, f"class{class_idx}", f"layer{layer_idx}")
            out[class_idx] = class_pipeline(model, batch, composite, layer_idx, class_idx, root, num_concepts=num_concepts,
                                            steps=steps, runs=runs, seed=seed, num_locations=num_locations, device=device,
                                            **optimizer_kwargs)
        return out
    from scipy.stats import ortho_group
    import time
    dev = torch.device(device)
    rank, world = dist.get_rank(), dist.get_world_size()
    classes = list(data_by_class.keys())

    def mark(name, t0):                                   # optional wall-clock breakdown (synchronises: diagnostics only)
        if timings is not None:
            torch.cuda.synchronize(dev)
            timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()
    t0 = time.perf_counter()
    # ---- stage 1 for every class on this rank's samples
    acts, ctxs = {}, {}
    for pos, class_idx in enumerate(classes):
        acts[pos], ctxs[pos] = pp.extract_context_pairs(model, data_by_class[class_idx], composite, layer_idx, class_idx,
                                                        num_locations=num_locations, normalize=True, device=dev)
    t0 = mark("stage1_all_classes", t0)
    mine = torch.tensor([acts[p].size(0) for p in range(len(classes))], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, len(classes), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    counts = counts.cpu()
    d = acts[0].size(-1)
    # ---- stage 2: whole classes side by side
    results: Dict[int, tuple] = {}
    plan = plan_class_schedule(len(classes), world)
    for round_plan in plan:                                       # every rank creates every group, in the same order
        for _, ranks in round_plan:
            if len(ranks) > 1:
                _subgroup(ranks)
    for round_plan in plan:
        t0 = time.perf_counter()
        act = redistribute_rows(acts, round_plan, counts, rank, world)
        ctx = redistribute_rows(ctxs, round_plan, counts, rank, world)
        t0 = mark("redistribute_rows", t0)
        my = next(((c, ranks) for c, ranks in round_plan if rank in ranks), None)
        if my is not None:
            pos, ranks = my
            class_idx = classes[pos]
            np.random.seed(seed)                                  # same start as class_pipeline / drsa.main (drsa.py:263-272)
            U = ortho_group.rvs(d)
            opt = None
            for run in range(1, runs + 1):
                U = U[:, np.random.permutation(d)]
                path = None
                if model_root is not None and rank == ranks[0]:
                    path = os.path.join(model_root, f"class{class_idx}", f"layer{layer_idx}", f"run{run}")
                    os.makedirs(path, exist_ok=True)
                opt = drsa.SubspaceOptimizer(torch.tensor(U, dtype=torch.float32), act, ctx, path, num_concepts=num_concepts,
                                             device=dev, process_group=_subgroup(ranks) if len(ranks) > 1 else False,
                                             **optimizer_kwargs)
                opt.run(steps=steps, save=path is not None)
            results[pos] = (opt.U, torch.as_tensor(opt.obj_history, device=dev), act.size(0))
            del opt
        del act, ctx
        mark(f"drsa_round_{len(round_plan)}_classes", t0)
    t0 = time.perf_counter()
    # ---- every rank ends with every class's U and objective history (d*m floats per class from the group's first rank)
    out = {}
    for round_plan in plan:
        for pos, ranks in round_plan:
            src = ranks[0]
            if rank == src:
                U, hist, rows = results[pos]
                meta = torch.tensor([hist.numel(), rows], dtype=torch.int64, device=dev)
            else:
                meta = torch.zeros(2, dtype=torch.int64, device=dev)
            dist.broadcast(meta, src)
            if rank != src:
                U = torch.empty(d, d, device=dev)
                hist = torch.empty(int(meta[0]), dtype=torch.float64, device=dev)
            hist = hist.to(torch.float64)
            dist.broadcast(U, src)
            dist.broadcast(hist, src)
            rows = results[pos][2] if pos in results else 0
            out[classes[pos]] = (U, hist.cpu().numpy(), rows)
    mark("broadcast_results", t0)
    return out
