"""Explanation-side consumers of the optimised subspaces -- mirror of the reference's
``cxai.xai.explain.explainer``: ``HeatmapGenerator`` (explainer.py:15-183), ``get_class_composite`` (:186-203) and
``compute_subspace_relevances`` (:206-242)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from drsa_audio_b200 import _lib as _L
from cxai.model.modify_model import ProjectionModel
from cxai.utils.constants import CLASS_IDX_MAPPER, CLASS_IDX_MAPPER_TOY
from cxai.xai.explain.attribute import SubspaceHook, compute_relevances
from cxai.xai.explain.rules import Epsilon, NameMapComposite, SequentialMergeBatchNorm

__all__ = ["HeatmapGenerator", "get_class_composite", "compute_subspace_relevances"]


class HeatmapGenerator:
    """Standard and concept-conditional relevance heatmaps in input space for one target class (explainer.py:15-183).
    Same constructor arguments, methods and ``info`` keys as the reference; the attribution runs on libdrsa_b200.so
    (one forward pass per sample, K+1 backward passes below the split layer -- the reference pushes K+1 clones of
    every sample through the whole network).

    ``canonizers``: extra keyword (not in the reference, whose GTZAN model for this class has no BatchNorm): pass
    ``[SequentialMergeBatchNorm()]`` for BatchNorm models."""

    def __init__(self, model, U: torch.Tensor, name_map, sample_class: str, num_concepts: int = 4, layer_idx: int = 10,
                 device=torch.device("cuda"), canonizers=()) -> None:
        self.device = torch.device(device) if isinstance(device, str) else device
        self.num_concepts = num_concepts
        case = "toy" if sample_class.endswith("1") or sample_class.endswith("2") else "gtzan"
        class_idx_mapper = CLASS_IDX_MAPPER if case == "gtzan" else CLASS_IDX_MAPPER_TOY
        self.class_idx = class_idx_mapper[sample_class]
        self.num_classes = len(class_idx_mapper)
        self.projectionmodel = ProjectionModel(model, layer_idx, U, self.num_concepts, case=case)
        self.composite = get_class_composite(name_map, self.num_concepts, device=device, canonizers=canonizers)
        self.info = {}

    def generate_subspace_heatmaps(self, input_batch: torch.Tensor, one_hot_encoded: bool = False,
                                   concept_flipping: bool = False, flip_all_classes: bool = False) -> None:
        """explainer.py:68-123: every sample is attributed K+1 times (standard + one per concept); results in
        ``self.info``: input, standard_heatmaps, standard_relevance, subspace_heatmaps (sorted by descending
        relevance per sample), subspace_relevances, mask."""
        self.info["input"] = input_batch.cpu().numpy()
        input_batch = input_batch.to(self.device)
        repeated_input_batch = input_batch.repeat_interleave(self.num_concepts + 1, dim=0)
        heatmaps = self.obtain_heatmaps(repeated_input_batch, one_hot_encoded, flip_all_classes).squeeze()
        heatmaps = heatmaps.view(-1, self.num_concepts + 1, heatmaps.size(-2), heatmaps.size(-1))
        heatmaps = heatmaps.detach().cpu().numpy()
        standard_heatmaps = heatmaps[:, 0:1]
        subspace_heatmaps = heatmaps[:, 1:]
        subspace_heatmaps, subspace_relevances, mask = self.sort_subspaces(subspace_heatmaps)
        self.info["standard_heatmaps"] = standard_heatmaps
        self.info["standard_relevance"] = standard_heatmaps.sum(axis=(-2, -1)).flatten()
        self.info["subspace_heatmaps"] = subspace_heatmaps
        self.info["subspace_relevances"] = subspace_relevances
        self.info["mask"] = mask

    def obtain_heatmaps(self, input_batch: torch.Tensor, one_hot_encoded: bool = False,
                        flip_all_classes: bool = False) -> torch.Tensor:
        """explainer.py:125-150.  (The reference evaluates ``len(self.num_classes)`` on an int there, SURVEY section 0;
        the intended value, the number of classes, is used.)"""
        return compute_relevances(self.projectionmodel, input_batch, self.composite, one_hot_encoded=one_hot_encoded,
                                  class_idx=self.class_idx if not flip_all_classes else None,
                                  num_classes=self.num_classes if flip_all_classes else None)

    def sort_subspaces(self, subspace_heatmaps: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """explainer.py:152-183: order the concepts of every sample by descending total relevance."""
        batch = subspace_heatmaps.shape[0]
        subspace_relevances = subspace_heatmaps.sum(axis=(-2, -1)).squeeze()
        if subspace_relevances.ndim == 1:                  # a single sample: the reference's squeeze() drops the batch axis
            subspace_relevances = subspace_relevances[None]
        mask = np.argsort(subspace_relevances, axis=-1)[..., ::-1]
        subspace_heatmaps = subspace_heatmaps[np.arange(batch)[:, None], mask]
        subspace_relevances = subspace_relevances[np.arange(batch)[:, None], mask]
        return subspace_heatmaps, subspace_relevances, mask


def get_class_composite(name_map: List[Tuple[List[str], object]], num_concepts: int, device=None, canonizers=()):
    """explainer.py:186-203: the given name map plus Epsilon on both projection layers and the SubspaceHook on the
    filter between them."""
    name_map_copy = list(name_map)
    name_map_copy.append((["features.invprojection"], Epsilon()))
    name_map_copy.append((["features.subspacefilter"], SubspaceHook(num_concepts, device=device)))
    name_map_copy.append((["features.projection"], Epsilon()))
    return NameMapComposite(name_map=name_map_copy, canonizers=canonizers)


def compute_subspace_relevances(act_vecs: torch.Tensor, ctx_vecs: torch.Tensor, U: torch.Tensor,
                                n_concepts: int = 4) -> torch.Tensor:
    """Per-instance concept relevances R_k = sum_p sum_{j in k} (a U)_j (c U)_j, no ReLU
    (explainer.py:206-242).  act/ctx: [batch, N, d] (or [N, d] for a single instance) -> [batch, K]."""
    assert act_vecs.dim() < 4 or ctx_vecs.dim() < 4, "Please provide act and ctx vectors reshaped to [batch, N, d]"
    if not torch.cuda.is_available():
        raise _L.DRSAError("no CUDA device available and no fallback path exists")
    a = act_vecs if act_vecs.dim() == 3 else act_vecs.unsqueeze(0)
    c = ctx_vecs if ctx_vecs.dim() == 3 else ctx_vecs.unsqueeze(0)
    dev = a.device if a.is_cuda else torch.device("cuda")
    a = a.detach().to(dev, torch.float32).contiguous()
    c = c.detach().to(dev, torch.float32).contiguous()
    Ud = U.detach().to(dev, torch.float32).contiguous()
    B, P, d = a.shape
    m = Ud.shape[1]
    lib = _L.lib()
    out = torch.empty(B, n_concepts, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(_L.check(lib.drsa_subspace_relevances_workspace_bytes(B, P, d, m))), dtype=torch.uint8,
                         device=dev)
        _L.check(lib.drsa_subspace_relevances(a.data_ptr(), c.data_ptr(), Ud.data_ptr(), B, P, d, m, n_concepts,
                                              out.data_ptr(), ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream().cuda_stream), "drsa_subspace_relevances")
    return out
