"""Prototype search -- mirror of the reference's cxai/xai/drsa/prototypes.py (``get_prototypes_ts`` :14-130): the
subset of ``n`` samples whose (activation, context) vectors reach the highest DRSA objective under a given U.

The reference loads the class's spectrograms from disk (``get_songs_drsa``, outside the path) and then, PER SUBSET, runs
an LRP pass and one ``obj_val``.  Here the LRP pass runs once over all N samples (minibatched on the device) and the
subsets only differ in which rows enter the objective: all their objectives come out of ONE batched library call
(``drsa_subset_objectives``: two projections of all rows + a segmented reduction) and one device->host copy."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from drsa_audio_b200 import _lib as _L
from cxai.xai.drsa import preprocessing as pp

__all__ = ["get_prototypes", "get_prototypes_ts", "subset_objectives"]


def subset_objectives(act: torch.Tensor, ctx: torch.Tensor, U: torch.Tensor, num_subsets: int, rows_per_subset: int,
                      num_concepts: int, max_rows: int = 1 << 21) -> torch.Tensor:
    """``obj_val`` (drsa.py:123-155) of ``num_subsets`` subsets of ``rows_per_subset`` consecutive rows each -> [num_subsets]
    on the device, in one library call per ``max_rows`` rows (drsa_subset_objectives) instead of one ``obj_val`` and one
    host synchronisation per subset (prototypes.py:98-119)."""
    lib = _L.lib()
    d, m = int(U.size(0)), int(U.size(1))
    out = torch.empty(num_subsets, dtype=torch.float32, device=act.device)
    sumsq = torch.empty(num_subsets, num_concepts, dtype=torch.float32, device=act.device)
    per_call = max(1, max_rows // rows_per_subset)
    with torch.cuda.device(act.device):
        ws = torch.empty(int(_L.check(lib.drsa_subspace_relevances_workspace_bytes(min(per_call, num_subsets), rows_per_subset,
                                                                                   d, m))), dtype=torch.uint8, device=act.device)
        s = torch.cuda.current_stream().cuda_stream
        for s0 in range(0, num_subsets, per_call):
            ns = min(per_call, num_subsets - s0)
            r0 = s0 * rows_per_subset
            _L.check(lib.drsa_subset_objectives(act[r0:].data_ptr(), ctx[r0:].data_ptr(), U.data_ptr(), ns, rows_per_subset, d,
                                                m, num_concepts, out[s0:].data_ptr(), sumsq[s0:].data_ptr(), ws.data_ptr(),
                                                ws.numel(), s), "drsa_subset_objectives")
    return out


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
    Ud = U.to(dev, torch.float32).contiguous()
    objs = subset_objectives(act, ctx, Ud, N // n, n * P, num_concepts)
    objs_host = objs.cpu()                              # the only synchronisation of the search
    best = (float(objs_host.max()), int(objs_host.argmax()))
    objs = objs_host.tolist()
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
