"""Stage 1 of the hot path -- drop-in for the reference's ``cxai.xai.drsa.preprocessing``.

Reference: cxai/xai/drsa/preprocessing.py (``preprocess_data`` :18, ``get_intermediate`` :106,
``compute_context_vectors`` :179, ``sample_spatial_locations`` :196, ``normalize_vectors`` :219,
``get_vectors_from_maps`` :234).  The disk/audio loaders of the reference file (:319-370) are out of
scope.  All device work goes through ``libdrsa_b200.so``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from drsa_audio_b200 import _lib as _L

__all__ = ["preprocess_data", "get_intermediate", "compute_context_vectors", "sample_spatial_locations",
           "normalize_vectors", "get_vectors_from_maps", "gather_context_pairs", "extract_context_pairs"]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda_f32(t: torch.Tensor) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise _L.DRSAError("no CUDA device available and no fallback path exists")
    if not t.is_cuda:
        t = t.cuda()
    return t.detach().to(torch.float32).contiguous()


def compute_context_vectors(activation_vectors: torch.Tensor, relevance_vectors: torch.Tensor) -> torch.Tensor:
    """c = R / (a + 1e-7) (preprocessing.py:179-193)."""
    a, R = _cuda_f32(activation_vectors), _cuda_f32(relevance_vectors)
    assert a.shape == R.shape
    out = torch.empty_like(a)
    if a.numel():
        with torch.cuda.device(a.device):
            _L.check(_L.lib().drsa_context_vectors(_ptr(a), _ptr(R), a.numel(), _ptr(out), _stream()),
                     "drsa_context_vectors")
    return out


def sample_spatial_locations(batch_size: int, map_size: Tuple[int, int], num_locations: int) -> np.ndarray:
    """Per sample, ``num_locations`` distinct flat positions drawn with the global numpy RNG, in the
    reference's call order (preprocessing.py:196-216) so that seeded runs pick the same positions."""
    idcs_batch = np.zeros((batch_size, num_locations), dtype=int)
    for i in range(batch_size):
        idcs_batch[i, :] = np.random.choice(map_size[0] * map_size[1], num_locations, replace=False)
    return idcs_batch


def normalize_vectors(vectors: torch.Tensor, process_group=None) -> torch.Tensor:
    """v / sqrt(mean(v^2)) / d^0.25 with the mean over ALL elements (preprocessing.py:219-231).
    With torch.distributed initialised the statistic is all-reduced, so row shards are normalised
    by the global value (2 scalars over NCCL)."""
    v = _cuda_f32(vectors).clone()
    d = v.size(-1)
    count = torch.tensor([v.numel()], dtype=torch.int64, device=v.device)
    ss = torch.zeros(1, dtype=torch.float64, device=v.device)
    lib = _L.lib()
    with torch.cuda.device(v.device):
        if v.numel():
            _L.check(lib.drsa_sumsq(_ptr(v), v.numel(), _ptr(ss), _stream()), "drsa_sumsq")
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(ss, group=process_group)
            torch.distributed.all_reduce(count, group=process_group)
        if v.numel():
            _L.check(lib.drsa_normalize(_ptr(v), v.numel() // d, d, _ptr(ss), int(count.item()), _stream()),
                     "drsa_normalize")
    return v


def get_vectors_from_maps(maps: torch.Tensor, idcs_batch: np.ndarray, layout: str = "reference") -> torch.Tensor:
    """Channel vectors at the given flat positions, [batch*num_locations, d].

    ``layout='reference'`` reproduces preprocessing.py:234-256 bit for bit, including the extra
    transpose that scrambles rows there; ``layout='fixed'`` returns row (b, l) = the
    channel vector at position ``idcs_batch[b, l]``.  Pure indexing (no arithmetic)."""
    batch_size, d, _, _ = maps.size()
    flat = maps.reshape(batch_size, d, -1)
    vectors = flat[np.arange(batch_size)[:, None], :, idcs_batch]          # [B, L, d]
    if layout == "reference":
        return vectors.transpose(-2, -1).reshape(-1, d)
    if layout == "fixed":
        return vectors.reshape(-1, d)
    raise ValueError("layout must be 'reference' or 'fixed'")


def gather_context_pairs(activation_maps: torch.Tensor, relevance_maps: torch.Tensor,
                         idcs_batch: Optional[np.ndarray] = None, normalize: bool = True, process_group=None):
    """Fused tail of stage 1: gather [N,d,H,W] maps at the sampled (or all) positions into
    row-major [N*L, d] activation vectors, form c = R/(a+1e-7) in the same pass, accumulate both sums
    of squares, and normalise (one kernel + two scale kernels instead of gather, divide, two
    reductions and two scalings).  Row layout is the corrected one."""
    a, R = _cuda_f32(activation_maps), _cuda_f32(relevance_maps)
    N, d, H, W = a.shape
    HW = H * W
    idx_t = None
    L = HW
    if idcs_batch is not None:
        idx_t = torch.as_tensor(np.ascontiguousarray(idcs_batch), dtype=torch.int64).to(a.device).contiguous()
        L = idx_t.shape[1]
    act = torch.empty(N * L, d, dtype=torch.float32, device=a.device)
    ctx = torch.empty_like(act)
    ss = torch.zeros(2, dtype=torch.float64, device=a.device)
    lib = _L.lib()
    with torch.cuda.device(a.device):
        _L.check(lib.drsa_context_gather(_ptr(a), _ptr(R), N, d, HW, _ptr(idx_t), L, _ptr(act), _ptr(ctx), _ptr(ss),
                                         _stream()), "drsa_context_gather")
        if normalize:
            count = torch.tensor([act.numel()], dtype=torch.int64, device=a.device)
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                torch.distributed.all_reduce(ss, group=process_group)
                torch.distributed.all_reduce(count, group=process_group)
            c = int(count.item())
            _L.check(lib.drsa_normalize(_ptr(act), N * L, d, _ptr(ss[0:]), c, _stream()), "drsa_normalize")
            _L.check(lib.drsa_normalize(_ptr(ctx), N * L, d, _ptr(ss[1:]), c, _stream()), "drsa_normalize")
    return act, ctx


def extract_context_pairs(model, input_batch, composite, layer_idx: int, class_idx: int,
                          num_locations: Optional[int] = None, one_hot_encoded: bool = False, normalize: bool = True,
                          process_group=None, device="cuda"):
    """Stage 1 in one call: spectrograms -> (activation vectors, context vectors) [N*L, d] at ``model.features[layer_idx]``,
    ready for ``SubspaceOptimizer`` -- ``get_intermediate`` :106 + ``sample_spatial_locations`` :196 (same RNG call order) +
    ``get_vectors_from_maps`` :234 (corrected row layout) + ``compute_context_vectors`` :179 + ``normalize_vectors`` :219.
    Where the split layer lies inside the tensor-core stack the rows are written straight from its NHWC planes
    (``drsa_context_pairs_nhwc``); otherwise the maps are formed and gathered as in the separate functions."""
    from cxai.xai.explain.lrp_engine import lrp_context_pairs
    if isinstance(input_batch, np.ndarray):
        input_batch = torch.tensor(input_batch)
    input_batch = input_batch.to(device)
    layer = model.features[layer_idx]
    idcs = None
    if num_locations:
        # the map size is needed before the pass: one sample through the plain route tells it (cheap, and only here)
        a1, _ = get_intermediate(model, input_batch[:1], composite, layer, class_idx, one_hot_encoded=one_hot_encoded)
        idcs = sample_spatial_locations(input_batch.size(0), tuple(a1.shape[-2:]), num_locations)
    res = lrp_context_pairs(model, input_batch, composite, layer, class_idx, idcs, one_hot_encoded=one_hot_encoded)
    if res is None:
        a_maps, R_maps = get_intermediate(model, input_batch, composite, layer, class_idx, one_hot_encoded=one_hot_encoded)
        return gather_context_pairs(a_maps, R_maps, idcs, normalize=normalize, process_group=process_group)
    act, ctx, ss = res
    if normalize and act.numel():
        lib = _L.lib()
        count = torch.tensor([act.numel()], dtype=torch.int64, device=act.device)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(ss, group=process_group)
            torch.distributed.all_reduce(count, group=process_group)
        c = int(count.item())
        with torch.cuda.device(act.device):
            _L.check(lib.drsa_normalize(_ptr(act), act.size(0), act.size(1), _ptr(ss[0:]), c, _stream()), "drsa_normalize")
            _L.check(lib.drsa_normalize(_ptr(ctx), ctx.size(0), ctx.size(1), _ptr(ss[1:]), c, _stream()), "drsa_normalize")
    return act, ctx


def get_intermediate(model, input_batch, composite, layer, class_idx, attr_batch_size: int = 64,
                     one_hot_encoded: bool = False):
    """Activation and relevance maps at ``layer`` (preprocessing.py:106-176)."""
    from cxai.xai.explain.lrp_engine import lrp_intermediate
    return lrp_intermediate(model, input_batch, composite, layer, class_idx, attr_batch_size, one_hot_encoded)


def preprocess_data(model, input_batch, composite, layer_idx: int, class_idx: int,
                    num_locations: Optional[int] = None, one_hot_encoded: bool = False, device="cuda"):
    """(activation_vectors, context_vectors) for DRSA (preprocessing.py:18-89).  The reference function
    cannot run as shipped; this keeps its signature and intent: maps at
    ``model.features[layer_idx]``, then either ``num_locations`` sampled positions per sample or all
    positions, then c = R/(a+1e-7).  Rows are [N*L, d] (corrected layout)."""
    if isinstance(input_batch, np.ndarray):
        input_batch = torch.tensor(input_batch)
    input_batch = input_batch.to(device)
    layer = model.features[layer_idx]
    a_maps, R_maps = get_intermediate(model, input_batch, composite, layer, class_idx, one_hot_encoded=one_hot_encoded)
    idcs = None
    if num_locations:
        idcs = sample_spatial_locations(a_maps.size(0), tuple(a_maps.shape[-2:]), num_locations)
    return gather_context_pairs(a_maps, R_maps, idcs, normalize=False)
