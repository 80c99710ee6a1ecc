"""Prototype search -- mirror of the reference's cxai/xai/drsa/prototypes.py (``get_prototypes_ts`` :14-130): the
subset of ``n`` samples whose (activation, context) vectors reach the highest DRSA objective under a given U.

The reference loads the class's spectrograms from disk (``get_songs_drsa``, outside the path) and then, PER SUBSET, runs
an LRP pass and one ``obj_val``.  Here the LRP pass runs once over all N samples (minibatched on the device) and the
subsets only differ in which rows enter the objective."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from cxai.xai.drsa import preprocessing as pp
from cxai.xai.drsa.drsa import SubspaceOptimizer, objective_fn

__all__ = ["get_prototypes", "get_prototypes_ts"]


def get_prototypes(model, layer_idx: int, U: torch.Tensor, composite, data_batch: torch.Tensor, class_idx: int,
                   num_concepts: int = 4, n: int = 10, N: Optional[int] = None, seed: int = 42,
                   loaded_samples: Optional[List[str]] = None, device="cuda"):
    """prototypes.py:59-130 on an in-memory batch.  Returns (activation vectors, context vectors of the best subset
    [n*P, d], indices of its samples in ``data_batch``, objective of every subset, names of its samples or None)."""
    dev = torch.device(device)
    N = N if N else data_batch.size(0)
    local_gen = torch.Generator().manual_seed(seed)
    perm_mask = torch.randperm(data_batch.size(0), generator=local_gen)          # prototypes.py:74-76
    sel = perm_mask[:N]
    batch = data_batch[sel].to(dev)
    a_maps, R_maps = pp.get_intermediate(model, batch, composite, model.features[layer_idx], class_idx)
    act, ctx = pp.gather_context_pairs(a_maps, R_maps, None, normalize=False)    # all positions, c = R / (a + 1e-7)
    P = act.size(0) // batch.size(0)
    Ud = U.to(dev)
    d_c = Ud.size(1) // num_concepts
    objs, best = [], (-1.0, None)
    for i in range(N // n):
        rows = slice(i * n * P, (i + 1) * n * P)
        obj = float(SubspaceOptimizer.obj_val(act[rows], ctx[rows], Ud, objective_fn, num_concepts, d_c))
        objs.append(obj)
        if obj > best[0]:
            best = (obj, i)
    i = best[1]
    rows = slice(i * n * P, (i + 1) * n * P)
    idx = sel[i * n:(i + 1) * n]
    names = [loaded_samples[j] for j in idx.tolist()] if loaded_samples is not None else None
    return act[rows].clone(), ctx[rows].clone(), idx, objs, names


def get_prototypes_ts(model, layer_idx: int, U: torch.Tensor, composite, path_to_data: str, sample_class: str,
                      case: str = "gtzan", num_concepts: int = 4, n: int = 10, N: int = None, excluded_folds: int = None,
                      seed: int = 42, device="cuda", data_batch: Optional[torch.Tensor] = None,
                      loaded_samples: Optional[List[str]] = None) -> Tuple[torch.Tensor, torch.Tensor, List[str], torch.Tensor]:
    """The reference's signature.  Reading the class's audio from ``path_to_data`` (fold lists, decoding) is outside the
    hot path: pass the spectrograms as ``data_batch`` (and their names as ``loaded_samples``)."""
    from cxai.utils.constants import CLASS_IDX_MAPPER, CLASS_IDX_MAPPER_TOY
    if data_batch is None:
        raise NotImplementedError("get_prototypes_ts: loading audio from path_to_data is outside this path; "
                                  "pass data_batch=<spectrograms of the class> (and loaded_samples=<their names>)")
    class_idx = (CLASS_IDX_MAPPER if case == "gtzan" else CLASS_IDX_MAPPER_TOY)[sample_class]
    a, c, idx, _, names = get_prototypes(model, layer_idx, U, composite, data_batch, class_idx, num_concepts, n, N, seed,
                                         loaded_samples, device)
    return a, c, names, idx
