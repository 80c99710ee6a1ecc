"""Projection layers that route relevance through the optimised subspaces -- mirror of the reference's
cxai/model/modify_model.py (``ProjectionModel`` :4-59, ``SubspaceFilter`` :62-72, ``Projection`` :75-97,
``InvProjection`` :100-123).

The modules hold U and the layer order; ``forward`` is the plain torch definition (so a ProjectionModel is an ordinary
``nn.Module``), while the explanation hot path -- ``compute_relevances`` / ``HeatmapGenerator`` -- runs them on
``libdrsa_b200.so`` through ``cxai.xai.explain.lrp_engine`` (``lrp_subspace_project`` / ``lrp_subspace_filter``).
Differences from the reference: the flatten size is taken from the wrapped model instead of the hard-coded 2048 / 64
 and InvProjection does not assume square feature maps (the map shape is remembered by Projection).
"""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ["ProjectionModel", "SubspaceFilter", "Projection", "InvProjection"]


class ProjectionModel(nn.Module):
    """Copy of ``model`` with Projection -> SubspaceFilter -> InvProjection inserted after ``features[layer_idx]``."""

    def __init__(self, model: nn.Module, layer_idx: int, U: torch.Tensor, num_concepts: int, case: str = "gtzan") -> None:
        super().__init__()
        assert 0 < layer_idx < len(model.features), "layer_idx has to be in range 0 - len(model.features)"
        self.num_flat_features = getattr(model, "num_flat_features", 2048 if case == "gtzan" else 64)
        self.layer_idx, self.num_concepts = layer_idx, num_concepts
        self.features = nn.Sequential()
        proj = Projection(U, num_concepts)
        for idx, layer in enumerate(model.features.children()):
            if idx == layer_idx + 1:
                self.features.add_module("projection", proj)
                self.features.add_module("subspacefilter", SubspaceFilter())
                self.features.add_module("invprojection", InvProjection(U, num_concepts, proj))
            self.features.add_module(str(idx), layer)
        self.classifier = nn.Sequential()
        for idx, layer in enumerate(model.classifier.children()):
            self.classifier.add_module(str(idx), layer)
        self.train(model.training)

    def forward(self, x):
        x = self.features(x)
        return self.classifier(x.reshape(x.size(0), -1))


class SubspaceFilter(nn.Module):
    """Identity; the place where the SubspaceHook masks the relevance of clone k to concept k."""

    def forward(self, act_map: torch.Tensor) -> torch.Tensor:
        return act_map


class Projection(nn.Module):
    """a [b, d, H, W] -> h [b, H*W, num_concepts, d_k] = a U per position."""

    def __init__(self, U: torch.Tensor, num_concepts: int) -> None:
        super().__init__()
        self.U = U
        self.num_concepts = num_concepts
        self.d_k = U.size(1) // num_concepts
        self.map_shape = None

    def forward(self, act_map: torch.Tensor) -> torch.Tensor:
        self.map_shape = tuple(act_map.shape[-2:])
        act_vecs = act_map.view(act_map.size(0), act_map.size(1), -1).transpose(-2, -1).contiguous()
        h = torch.matmul(act_vecs, self.U.to(act_vecs))
        return h.view(h.size(0), h.size(1), self.num_concepts, self.d_k)


class InvProjection(nn.Module):
    """h [b, H*W, num_concepts, d_k] -> a' [b, d, H, W] = h U^T per position."""

    def __init__(self, U: torch.Tensor, num_concepts: int, projection: Projection = None) -> None:
        super().__init__()
        self.U_inv = U.T
        self.num_concepts = num_concepts
        self.d = self.U_inv.size(1)
        self.d_k = self.U_inv.size(0) // num_concepts
        self._projection = [projection]         # not a sub-module: only consulted for the map shape

    def forward(self, h: torch.Tensor) -> torch.Tensor:
        b, n, _, _ = h.size()
        a_ = torch.matmul(h.reshape(b, n, -1), self.U_inv.to(h))
        proj = self._projection[0]
        if proj is not None and proj.map_shape is not None and proj.map_shape[0] * proj.map_shape[1] == n:
            fh, fw = proj.map_shape
        else:
            fh = fw = int(n ** .5)                # the reference's assumption (modify_model.py:121)
        return a_.transpose(-2, -1).reshape(b, self.d, fh, fw).contiguous()
