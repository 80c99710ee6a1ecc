"""Per-class DRSA data sets on disk -- mirror of the reference's cxai/xai/drsa/cluster/getdrsadata.py
(``save_data`` :26-45, ``load_and_normalize_data`` :48-59; its ``main`` :63-139 is a hard-coded cluster script whose
loop is ``extract_class_datasets`` here).  File format: ``{output_path}/{case}/{model}/{sample_class}/dataset_layer{L}.pkl``
holding a pickled ``list`` of ``(activation_vector, context_vector)`` numpy pairs, one per row."""
from __future__ import annotations

import os
import pickle
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from cxai.xai.drsa.preprocessing import normalize_vectors, preprocess_data

__all__ = ["save_data", "load_and_normalize_data", "extract_class_datasets"]


def save_data(activation_vectors, context_vectors, layer=None, sample_class=None, case="gtzan", model="bn",
              output_path=None) -> str:
    assert type(layer) == int, "layer has to be defined and of type int"
    assert type(sample_class) == str, "sample_class has to be defined and of type str"
    assert output_path is not None, "please provide an output path to save the data"
    paired_dataset = list(zip(np.asarray(activation_vectors), np.asarray(context_vectors)))
    path = os.path.join(output_path, f"{case}/{model}/{sample_class}")
    os.makedirs(path, exist_ok=True)
    filepath = os.path.join(path, f"dataset_layer{layer}.pkl")
    with open(filepath, "wb") as file:
        pickle.dump(paired_dataset, file)
    return filepath


def load_and_normalize_data(filepath, device="cuda"):
    """(normalised activation vectors, normalised context vectors) on ``device`` (getdrsadata.py:48-59)."""
    with open(filepath, "rb") as file:
        dataset = pickle.load(file)
    a, c = zip(*dataset)
    a = torch.tensor(np.array(a), device=device).detach().requires_grad_(False)
    c = torch.tensor(np.array(c), device=device).detach().requires_grad_(False)
    return normalize_vectors(a), normalize_vectors(c)


def extract_class_datasets(model, data_by_class: Dict[str, torch.Tensor], composite, class_idx_mapper: Dict[str, int],
                           layer_idcs: Iterable[int], output_path: str, num_locations: Optional[int] = 20,
                           case: str = "gtzan", model_name: str = "bn", device="cuda"):
    """The loop of getdrsadata.main (:118-139): for every class and split layer run the LRP pass on that class's
    spectrograms, form (a, c) pairs at ``num_locations`` sampled positions and store them in the reference's format."""
    written = []
    for genre, batch in data_by_class.items():
        for layer_idx in layer_idcs:
            act, ctx = preprocess_data(model, batch, composite, layer_idx, device=device,
                                       class_idx=class_idx_mapper[genre], num_locations=num_locations)
            written.append(save_data(act.detach().cpu().numpy(), ctx.detach().cpu().numpy(), layer=layer_idx,
                                     sample_class=genre, case=case, model=model_name, output_path=output_path))
    return written
