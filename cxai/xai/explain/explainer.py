"""Explanation-side consumers of the optimised subspaces -- mirror of the reference's
``cxai.xai.explain.explainer`` for the rows of SURVEY section 8 that are on the hot path:
``compute_subspace_relevances`` (explainer.py:206-242).  ``HeatmapGenerator`` is a "next" row."""
from __future__ import annotations

import torch

from drsa_audio_b200 import _lib as _L

__all__ = ["compute_subspace_relevances"]


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
