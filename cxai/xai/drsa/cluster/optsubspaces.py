"""Per-class subspace optimisation drivers -- mirror of the reference's cxai/xai/drsa/cluster/optsubspaces.py
(:8-52: for every class and split layer load the stored pairs, normalise, ``drsa.main``), plus the BASELINE cfg-5
pipeline that never leaves the device: spectrograms -> CNN forward -> LRP to the split layer -> (a, c) pairs ->
normalise -> DRSA, once per class.  With ``torch.distributed`` initialised every rank passes ITS samples of the class:
the LRP pass needs no communication, the normalisation statistics and the d*m + K row sums of every DRSA step are
all-reduced (rows never move), and every rank ends with the same U."""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from cxai.xai.drsa import drsa
from cxai.xai.drsa import preprocessing as pp
from cxai.xai.drsa.cluster.getdrsadata import load_and_normalize_data

__all__ = ["optimize_stored_classes", "class_pipeline", "all_classes_pipeline"]


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
    a_maps, R_maps = pp.get_intermediate(model, input_batch.to(dev), composite, model.features[layer_idx], class_idx)
    idcs = None
    if num_locations:
        idcs = pp.sample_spatial_locations(a_maps.size(0), tuple(a_maps.shape[-2:]), num_locations)
    act, ctx = pp.gather_context_pairs(a_maps, R_maps, idcs, normalize=True)
    del a_maps, R_maps
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


def all_classes_pipeline(model, data_by_class: Dict[int, torch.Tensor], composite, layer_idx: int, model_root: Optional[str],
                         num_concepts: int = 4, steps: int = 2000, runs: int = 1, seed: int = 42,
                         num_locations: Optional[int] = None, device="cuda", **optimizer_kwargs):
    """BASELINE cfg 5: the per-class pipeline for every class (class index -> this rank's spectrograms of that class)."""
    out = {}
    for class_idx, batch in data_by_class.items():
        root = None if model_root is None else os.path.join(model_root, f"class{class_idx}", f"layer{layer_idx}")
        out[class_idx] = class_pipeline(model, batch, composite, layer_idx, class_idx, root, num_concepts=num_concepts,
                                        steps=steps, runs=runs, seed=seed, num_locations=num_locations, device=device,
                                        **optimizer_kwargs)
    return out
