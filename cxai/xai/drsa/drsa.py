"""DRSA subspace optimisation on B200 -- drop-in for the reference's ``cxai.xai.drsa.drsa``.

Same names, arguments and on-disk outputs as the reference module
(``cxai/xai/drsa/drsa.py`` in sharckhai/drsa-audio: ``SubspaceOptimizer`` :15-168,
``generalized_fmean`` :171, ``project_grad`` :186, ``orthogonalize`` :202,
``objective_fn`` :224, ``main`` :241), but every device operation on the path is a call
into ``libdrsa_b200.so`` (hand-written sm_100a CUDA, see ``include/drsa_b200.h``):

* one fused row pass per step (projection, per-concept ReLU'd relevance, pooling sums and the
  un-normalised gradient) instead of 4 GEMMs + ~10 elementwise kernels + autograd,
* the ascent step and the polar retraction on the device (the reference round-trips through
  a host fp64 ``eigh`` each step, drsa.py:216),
* the objective history kept in a device buffer and copied once at the end (the reference
  synchronises twice per step, drsa.py:104 and :216).

Rows shard across ranks: with ``torch.distributed`` initialised (NCCL), each rank holds a
slice of the rows, U is replicated, and one all-reduce of ``d*m + K`` floats per step joins
the row sums.  There is no CPU / PyTorch fallback: without the CUDA library the calls raise.
"""
from __future__ import annotations

import math
import os
import pickle
from typing import List, Optional

import numpy as np
import torch

from drsa_audio_b200 import _lib as _L

__all__ = ["SubspaceOptimizer", "generalized_fmean", "project_grad", "orthogonalize", "objective_fn", "main"]

_DEFAULT_DEVICE = torch.device("cuda")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _as_device(device) -> torch.device:
    dev = device if isinstance(device, torch.device) else torch.device(device)
    if dev.type != "cuda":
        raise _L.DRSAError(f"this build of cxai.xai.drsa runs on CUDA (sm_100a) only, got device '{dev}'")
    if not torch.cuda.is_available():
        raise _L.DRSAError("no CUDA device available and no fallback path exists")
    return dev


def _f32c(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _pow2_scale(absmax: float) -> float:
    """Power of two s with absmax*s in [128, 256): keeps fp16 far from overflow and lifts small
    values out of the subnormal range; exact to undo."""
    if not math.isfinite(absmax) or absmax <= 0.0:
        return 1.0
    return float(2.0 ** (7 - math.floor(math.log2(absmax))))


_PRECISIONS = {"fp32": _L.PREC_FP32, "tc": _L.PREC_TC_F16, "tc_split": _L.PREC_TC_F16X2, "tc_hilo": _L.PREC_TC_F16_AC2,
               "tc32": _L.PREC_TC_F32C, "tc_dc": _L.PREC_TC_F16_AC2}
# 'tc_dc': steps between two evaluations of the row-rounding correction.  Measured final angles over 2 000 steps against the
# reference's golden trajectories: every 16 steps 1.3e-4 .. 2.4e-4 rad from 262 144 rows on, but 5.7e-4 .. 7.2e-4 rad at 8 192
# rows, where every 8 steps gives 3.1e-4 rad -- small problems correct twice as often.
DC_EVERY = 16
DC_EVERY_SMALL = 8
DC_EVERY_LARGE = 32          # from 524 288 rows: the correction itself shrinks like 1/sqrt(M) (simulation at cfg 2: 4.2e-5 rad at
                             # n = 8, 4.8e-5 at 16, 5.8e-5 at 32; measured by bench.py's parity block)


def dc_every(M_global: int) -> int:
    if M_global >= 2 * _AUTO_DC_FROM:
        return DC_EVERY_LARGE
    return DC_EVERY if M_global >= _AUTO_DC_FROM else DC_EVERY_SMALL
_PACK_CHUNK_ROWS = 1 << 20          # rows per host->device chunk when the fp32 rows do not fit next to the packed copy


class _RowPass:
    """Device state of one row shard: the (packed) rows, workspaces and the step/finish calls.

    ``act`` / ``ctx`` may live on the host: the tensor-core modes only ever need the packed fp16 planes on the device, so
    the fp32 rows are uploaded (whole if they fit, else in chunks of ``_PACK_CHUNK_ROWS`` rows), packed and released --
    cfg 4 (12.8 M rows x 512 channels = 26 GB per fp32 matrix) packs to 13 GB per matrix ('tc')."""

    def __init__(self, act: torch.Tensor, ctx: torch.Tensor, d: int, m: int, K: int, precision: str, dev: torch.device):
        lib = _L.lib()
        self.lib = lib
        self.M, self.d, self.m, self.K = int(act.shape[0]), d, m, K
        if precision not in _PRECISIONS:
            raise ValueError("precision must be 'auto', 'fp32', 'tc', 'tc_split', 'tc_hilo', 'tc_dc' or 'tc32'")
        self.precision = precision
        self.prec_code = _PRECISIONS[precision]
        if lib.drsa_step_workspace_bytes(max(self.M, 1), d, m, K, self.prec_code) < 0:
            raise _L.DRSAError(f"precision='{precision}' does not support d={d}, m={m}, K={K}")
        self.is_tc = precision != "fp32"
        self.u_split = precision in ("tc_split", "tc32")
        self.hilo = precision in ("tc_hilo", "tc32", "tc_dc")
        # U rounded to fp16 once per step, with error feedback through the lo plane (drsa_finish_step, u_rounded = 2)
        self.u_rounded = 2 if precision in ("tc", "tc_hilo", "tc_dc") else 0
        self.dc = precision == "tc_dc"
        self.sums = torch.zeros(d * m + K, dtype=torch.float32, device=dev)
        if self.dc:                 # single-plane sums of the current step, and the correction kept between evaluations
            self.sums_raw = torch.zeros_like(self.sums)
            self.delta = torch.zeros_like(self.sums)
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)
        self.scaleA = self.scaleC = self.pq_scale = 1.0
        self.Ut_hi = self.Ut_lo = None
        if self.M > 0:
            ws = _L.check(lib.drsa_step_workspace_bytes(self.M, d, m, K, self.prec_code), "drsa_step_workspace_bytes")
            self.ws_step = torch.empty(max(int(ws), 256), dtype=torch.uint8, device=dev)
        if self.is_tc:
            self.Ut_hi = torch.empty(m, d, dtype=torch.float16, device=dev)
            if self.u_split or self.u_rounded == 2:
                self.Ut_lo = torch.empty(m, d, dtype=torch.float16, device=dev)
            self.A, self.scaleA, rhoA = self._pack(act, dev)
            self.C, self.scaleC, rhoC = self._pack(ctx, dev)
            # |g*HC| <= pq * rhoA * rhoC^2 and |g*HA| <= pq * rhoA^2 * rhoC (rho = largest packed row norm):
            # choose the power of two pq that keeps both below 2^15, so fp16 P/Q can never overflow
            bound = max(rhoA * rhoC * rhoC, rhoA * rhoA * rhoC, 1e-30)
            self.pq_scale = float(2.0 ** math.floor(math.log2(32768.0 / bound)))
        else:
            self.A, self.C = _f32c(act, dev), _f32c(ctx, dev)
        wf = _L.check(lib.drsa_finish_workspace_bytes(d, m), "drsa_finish_workspace_bytes")
        self.ws_fin = torch.empty(int(wf), dtype=torch.uint8, device=dev)

    def _pack(self, x: torch.Tensor, dev: torch.device):
        """fp16 copy of the rows (one plane, or hi + lo planes back to back), its power-of-two scale and the largest
        packed row norm."""
        M, d = int(x.shape[0]), int(x.shape[1])
        out = torch.empty((2, M, d) if self.hilo else (M, d), dtype=torch.float16, device=dev)
        if M == 0:
            return out, 1.0, 0.0
        x = x.detach()
        if x.is_cuda or 2 * x.numel() * 4 < torch.cuda.mem_get_info(dev)[0]:
            chunks = [(0, M)]
            whole = _f32c(x, dev)
            get = lambda r0, r1: whole
        else:                                   # two passes over the host rows (statistics, then packing)
            chunks = [(r0, min(r0 + _PACK_CHUNK_ROWS, M)) for r0 in range(0, M, _PACK_CHUNK_ROWS)]
            get = lambda r0, r1: _f32c(x[r0:r1], dev)
        stats = torch.zeros(2, dtype=torch.float32, device=dev)
        tmp = torch.zeros(2, dtype=torch.float32, device=dev)
        for r0, r1 in chunks:
            xc = get(r0, r1)
            _L.check(self.lib.drsa_absmax(_ptr(xc), xc.numel(), _ptr(tmp[0:]), _stream()), "drsa_absmax")
            _L.check(self.lib.drsa_rownorm_max(_ptr(xc), xc.size(0), d, _ptr(tmp[1:]), _stream()), "drsa_rownorm_max")
            torch.maximum(stats, tmp, out=stats)
        mx, rho = (float(v) for v in stats.cpu())          # one host sync, once per optimiser
        scale = _pow2_scale(mx)
        flat = out.view(2, -1) if self.hilo else out.view(1, -1)
        for r0, r1 in chunks:
            xc = get(r0, r1)
            hi = flat[0, r0 * d:]
            if self.hilo:
                _L.check(self.lib.drsa_pack_f16_hilo(_ptr(xc), xc.numel(), scale, _ptr(hi), _ptr(flat[1, r0 * d:]),
                                                     _stream()), "drsa_pack_f16_hilo")
            else:
                _L.check(self.lib.drsa_pack_f16(_ptr(xc), xc.numel(), scale, _ptr(hi), _stream()), "drsa_pack_f16")
        return out, scale, rho * scale

    def split_u(self, U: torch.Tensor):
        if self.is_tc:
            _L.check(self.lib.drsa_split_u(_ptr(U), self.d, self.m, _ptr(self.Ut_hi), _ptr(self.Ut_lo), _stream()),
                     "drsa_split_u")

    def _row_sums(self, U: torch.Tensor, code: int, out: torch.Tensor):
        _L.check(self.lib.drsa_step(_ptr(self.A), _ptr(self.C), _ptr(U), _ptr(self.Ut_hi), _ptr(self.Ut_lo), self.M,
                                    self.d, self.m, self.K, code, self.scaleA, self.scaleC, self.pq_scale,
                                    _ptr(out), _ptr(self.ws_step), self.ws_step.numel(), _stream()), "drsa_step")

    def step(self, U: torch.Tensor, correct: bool = True):
        """Row sums of this shard into self.sums (drsa.py:148-155 + backward of :100).

        'tc_dc' (deferred correction): a step with ``correct`` evaluates the sums twice at the same U -- on the hi planes of
        the rows alone and on hi + lo -- uses the latter and keeps the difference; the other steps read only the hi planes
        (the first M*d elements of the packed buffer are a valid single-plane matrix) and add the kept difference."""
        if self.M == 0:
            self.sums.zero_()
            return
        if not self.dc:
            self._row_sums(U, self.prec_code, self.sums)
            return
        n = self.sums.numel()
        self._row_sums(U, _L.PREC_TC_F16, self.sums_raw)
        if correct:
            self._row_sums(U, _L.PREC_TC_F16_AC2, self.sums)
            _L.check(self.lib.drsa_sums_combine(_ptr(self.sums), _ptr(self.sums_raw), -1.0, _ptr(self.delta), n, _stream()),
                     "drsa_sums_combine")
        else:
            _L.check(self.lib.drsa_sums_combine(_ptr(self.sums_raw), _ptr(self.delta), 1.0, _ptr(self.sums), n, _stream()),
                     "drsa_sums_combine")

    def finish(self, U: torch.Tensor, M_global: int, obj_log: Optional[torch.Tensor], log_index: int, update: bool,
               max_iters: int, tol: float, px: "Optional[_PeerExchange]" = None):
        # with error feedback (u_rounded = 2) the stored fp16 matrix is an INPUT as well: it is what the row pass used
        hi = _ptr(self.Ut_hi) if (update or self.u_rounded == 2) else None
        lo = _ptr(self.Ut_lo) if update else None
        if px is not None:          # all-reduce of self.sums over the ranks inside the kernel (peer memory)
            import ctypes as C
            _L.check(self.lib.drsa_finish_step_p2p(C.byref(px.desc), _ptr(self.sums), M_global, _ptr(U), self.d, self.m,
                                                   self.K, _ptr(U) if update else None, hi, lo, _ptr(obj_log), log_index,
                                                   max_iters, tol, self.u_rounded, _ptr(self.status), _ptr(self.ws_fin),
                                                   self.ws_fin.numel(), _stream()), "drsa_finish_step_p2p")
            return
        _L.check(self.lib.drsa_finish_step(_ptr(self.sums), M_global, _ptr(U), self.d, self.m, self.K,
                                           _ptr(U) if update else None, hi, lo, _ptr(obj_log), log_index, max_iters, tol,
                                           self.u_rounded, _ptr(self.status), _ptr(self.ws_fin), self.ws_fin.numel(),
                                           _stream()), "drsa_finish_step")


class _PeerExchange:
    """Exchange buffers for the all-reduce that is fused into the finish kernel (drsa_finish_step_p2p): one buffer per
    rank, every rank maps all of them (NVLink / NVSwitch peer memory).  Buffers are shared through cudaIpc handles sent
    over the process group ('ipc'), or come from torch's symmetric memory ('symm').  They are cached per
    (group, device, size) and live until the process exits: the protocol state inside them (exchange counter, parity)
    stays consistent as long as every rank makes the same sequence of calls, so optimisers can share them."""

    _cache = {}

    def __init__(self, world: int, rank: int, nbytes: int, ptrs, keep=None):
        self.world, self.rank, self.nbytes = world, rank, nbytes
        self.desc = _L.PeerExchange()
        self.desc.world, self.desc.rank = world, rank
        for r in range(world):
            self.desc.buffers[r] = ptrs[r]
        self._keep = keep              # symmetric-memory tensor / handle: must outlive the exchange

    @staticmethod
    def _vote(ok: bool, group, dev) -> bool:
        """Joint decision: True only if every rank of the group succeeded so far.  Every rank calls this the same number
        of times, so a rank that failed never leaves the others alone inside a collective."""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=group)
        return int(flag.item()) == 1

    @classmethod
    def _setup(cls, group, world: int, rank: int, nbytes: int, dev: torch.device, how: str):
        """Collective.  Phases with a vote after each: (1) allocate this rank's buffer, (2) exchange the handles,
        (3) map the peers' buffers.  A failed vote releases what this rank holds and returns None on EVERY rank."""
        import ctypes as C
        import warnings
        lib = _L.lib()
        dist = torch.distributed
        if how == "symm":
            t = hdl = None
            try:
                import torch.distributed._symmetric_memory as symm
                t = symm.empty(nbytes // 4, dtype=torch.float32, device=dev)
                t.zero_()
                err = None
            except Exception as e:          # noqa: BLE001
                err = e
            if not cls._vote(err is None, group, dev):
                if err is not None:
                    warnings.warn(f"DRSA: symmetric memory unavailable ({err}); using the NCCL all-reduce")
                return None
            try:                             # rendezvous is itself a collective: every rank reaches it (vote above)
                hdl = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                err = None
            except Exception as e:          # noqa: BLE001
                err = e
            if not cls._vote(err is None, group, dev):
                return None
            return cls(world, rank, nbytes, ptrs, keep=(t, hdl))
        # ---- cudaIpc
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        st = lib.drsa_ipc_alloc(nbytes, C.byref(own), handle)
        if not cls._vote(st == 0, group, dev):
            if st == 0:
                lib.drsa_ipc_free(own)
            else:
                warnings.warn(f"DRSA: peer buffer allocation failed ({_L.status_string(st)}); using the NCCL all-reduce")
            return None
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        ptrs, opened, ok = [], [], True
        for r in range(world):
            if r == rank:
                ptrs.append(own.value)
                continue
            q = C.c_void_p()
            st = lib.drsa_ipc_open(handles[r], C.byref(q))
            if st != 0:
                ok = False
                warnings.warn(f"DRSA: cannot map the buffer of rank {r} ({_L.status_string(st)}); using the NCCL all-reduce")
                break
            opened.append(q)
            ptrs.append(q.value)
        if not cls._vote(ok, group, dev):
            for q in opened:                 # close every mapping before the owners free (drsa_b200.h)
                lib.drsa_ipc_close(q)
            dist.barrier(group=group)
            lib.drsa_ipc_free(own)
            return None
        return cls(world, rank, nbytes, ptrs)

    @classmethod
    def get(cls, group, nbytes: int, dev: torch.device, how: str):
        """Collective over ``group``: returns the shared exchange (or None on every rank if any rank failed)."""
        dist = torch.distributed
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        # one exchange per stream: the protocol state inside the buffers assumes stream-ordered calls
        key = (id(group) if group is not None else 0, str(dev), nbytes, how, torch.cuda.current_stream(dev).cuda_stream)
        px = cls._cache.get(key)
        # the cache is filled collectively, so either every rank has the entry or none has
        if px is None:
            px = cls._setup(group, world, rank, nbytes, dev, how)
            if px is None:
                return None
            cls._cache[key] = px
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)       # every buffer is zero-filled and mapped before anyone pushes into it
        return px


# Long-horizon accuracy of the arithmetic modes (tests/test_gpu_drsa_long.py, scripts/horizon_parity.py,
# scripts/sim_feedback_gpu.py).  Over the reference's horizon of 2 000 steps (drsa.py:76) two roundings move the final
# subspaces against the fp32 trajectory (tolerance 1e-3 rad):
#   * the per-step rounding of U to fp16 (3.8e-4 .. 7e-3 rad, worst where the objective has flat directions) -- removed by
#     the error feedback of drsa_finish_step (u_rounded = 2): 1e-5 .. 1.4e-4 rad, at no cost;
#   * storing the rows once in fp16, a FIXED perturbation of the data set: 4.4e-4 rad at M = 640 000 (d = 256) but
#     1e-3 .. 5.7e-3 rad at M = 8 192 .. 65 536 -- removed by hi + lo planes ('tc_hilo', 2x the MMA work: 2e-5 .. 1.3e-4 rad
#     measured) or, at (n+2)/n of the single-plane cost, by the deferred correction 'tc_dc' (5e-5 .. 3e-4 rad at n = 8).
# 'auto': CUDA-core fp32 below 8 192 rows per rank (the tensor cores do not matter there), 'tc_hilo' for medium problems (the
# step is latency-bound anyway), 'tc_dc' from 262 144 rows.  'tc' (single plane, no correction) is never chosen
# automatically: it is the fastest mode and within tolerance at cfg 2, but with only a factor of two in hand.
_AUTO_FP32_BELOW = 8192
_AUTO_DC_FROM = 262_144


def _auto_precision(M_global: int, avg_rows: int, d: int, m: int, K: int) -> str:
    if avg_rows < _AUTO_FP32_BELOW:
        return "fp32"
    return "tc_dc" if M_global >= _AUTO_DC_FROM else "tc_hilo"


def _pad_plan(d: int, m: int, K: int, rows: int):
    """Shape (d', m') of an equivalent zero-padded problem that the tensor-core row pass covers, or None.

    The reference's production split layers are not all tensor-core shapes (arch A, getdrsadata.py:72-73: layer 19 has
    d = 100, i.e. K = 4 concepts of d_k = 25).  Padding is exact: every concept block is widened to d_k' in {32, 64, 128}
    columns, the m' - m new columns are unit vectors in d' - d new (all-zero) channels of A and C, so projections,
    relevances, pooling sums and the gradient of the original entries are unchanged, the new entries have zero gradient,
    and the polar factor of the block-diagonal matrix is block-diagonal."""
    lib = _L.lib()
    if m % K != 0:
        return None
    d_k = m // K
    for dkp in (32, 64, 128):
        if dkp < d_k:
            continue
        mp = K * dkp
        if mp % 128 != 0:
            continue
        for dp in (128, 256, 512):
            if dp >= mp and dp - d >= mp - m and dp >= d and \
                    lib.drsa_step_workspace_bytes(max(rows, 1), dp, mp, K, _L.PREC_TC_F16) >= 0:
                return dp, mp, dkp
    return None


class SubspaceOptimizer:
    """Gradient ascent on the DRSA objective with a polar retraction after every step.

    Mirrors the reference class (drsa.py:15-168): same constructor arguments and attributes
    (``U, act_vecs, ctx_vecs, num_concepts, d_k, obj_fn, device, path_to_model``), ``run``,
    ``obj_val``, ``save_model`` and ``save_train_stats``.

    Extra keyword arguments (all optional, defaults keep the reference behaviour):
        precision: 'auto' | 'fp32' | 'tc' | 'tc_split' | 'tc_hilo' | 'tc32' -- arithmetic of the row pass (see
            include/drsa_b200.h): 'fp32' = CUDA cores; 'tc' = tcgen05, rows stored once in fp16, U rounded to fp16 once
            per step with a first-order corrected objective; 'tc_split' = 'tc' with U split hi + lo; 'tc_hilo' = 'tc' with
            the rows stored as hi + lo fp16 planes (22 bits, 2x the MMA work); 'tc_dc' = 'tc' plus a deferred correction
            of the row rounding evaluated every ``DC_EVERY`` steps ((n+2)/n of the cost of 'tc'); 'tc32' = rows AND U
            hi + lo: fp32-class operands on the tensor cores (2.6x).  In 'tc', 'tc_hilo' and 'tc_dc' the per-step rounding
            of U uses error feedback.  'auto' picks by row count so that the result stays within 1e-3 rad of the fp32
            trajectory over the reference's 2 000 steps (see ``_auto_precision``).
        process_group: torch.distributed group over which the rows are sharded; ``activation_vecs``
            / ``context_vecs`` are then this rank's slice.  Defaults to the world group if
            torch.distributed is initialised; ``False`` = this rank alone (no exchange).
        retraction_iters / retraction_tol: bounds of the on-device Newton-Schulz polar iteration (it stops as soon
            as it has converged: 4-5 sweeps in a normal step; the first steps of a tiny problem, where the
            gradient dwarfs U, need 15-25).
        use_cuda_graph: capture one step and replay it (single process, or several ranks with the peer exchange).
        retraction: 'polar' (default: the reference's retraction) or 'qr' (Q of the QR factorisation with positive diagonal,
            for comparison only: a different trajectory; fp32, single rank).
        exchange: how the ``d*m + K`` row sums are joined across ranks: 'p2p' = inside the finish kernel over
            NVLink peer memory (drsa_finish_step_p2p; buffers shared through cudaIpc), 'p2p_symm' = the same with
            torch symmetric memory, 'nccl' = ``all_reduce`` between the two kernels, 'auto' = 'p2p' when the process
            group runs on NCCL with at most 8 ranks and the shape is supported, else 'nccl'.
    """

    def __init__(self, U: torch.Tensor, activation_vecs: torch.Tensor, context_vecs: torch.Tensor,
                 path_to_model: Optional[str], num_concepts: int = 4, device=_DEFAULT_DEVICE, *,
                 precision: str = "auto", process_group=None, retraction_iters: int = 40,
                 retraction_tol: float = 1e-6, use_cuda_graph: bool = True, exchange: str = "auto",
                 retraction: str = "polar") -> None:
        assert num_concepts > 0, "num_concepts must be a positive number"
        assert U.size(1) % num_concepts == 0, "num_concepts must be a divisor of the number of columns of U"
        assert activation_vecs.shape == context_vecs.shape and activation_vecs.dim() == 2
        assert activation_vecs.size(1) == U.size(0), "vector dimension must equal the number of rows of U"
        if retraction not in ("polar", "qr"):
            raise ValueError("retraction must be 'polar' (the reference's, drsa.py:201-221) or 'qr'")
        if retraction == "qr":
            # comparison only (BASELINE north_star (3)): Q of QR(U + grad) with diag(R) > 0 is NOT the reference's
            # retraction and the trajectories differ; fp32 arithmetic, single rank
            if precision not in ("auto", "fp32"):
                raise _L.DRSAError("retraction='qr' runs in fp32 arithmetic only")
            precision, process_group = "fp32", False
        self.retraction = retraction
        self.device = _as_device(device)
        self.path_to_model = path_to_model
        self.num_concepts = num_concepts
        self.d_k = U.size(1) // num_concepts          # drsa.py:68 (U square there)
        self.obj_fn = objective_fn
        d, m = int(U.size(0)), int(U.size(1))
        self._pad = None
        with torch.cuda.device(self.device):
            self._Uw = _f32c(U, self.device).clone()          # the matrix the kernels work on (padded if self._pad)
            # the rows as the caller passed them (host or device); a device fp32 copy is only made where the arithmetic
            # needs one ('fp32') or on access of .act_vecs / .ctx_vecs -- the tensor-core modes keep only the packed planes
            self._act_src, self._ctx_src = activation_vecs.detach(), context_vecs.detach()
            self._act_dev = self._ctx_dev = None
            # process_group=False: this rank optimises its rows ALONE although torch.distributed is initialised (independent
            # problems scheduled across the GPUs, e.g. one class per GPU in the cfg-5 pipeline)
            self._dist = process_group is not False and torch.distributed.is_available() and \
                torch.distributed.is_initialized()
            self._group = process_group if process_group is not False else None
            rows = int(activation_vecs.size(0))
            M_local = torch.tensor([rows], dtype=torch.int64, device=self.device)
            if self._dist:
                torch.distributed.all_reduce(M_local, group=self._group)
            self.M_global = int(M_local.item())
            # every rank must take the same decisions (arithmetic, padded shapes of the exchange): they are derived from
            # the average shard size and the global row count, not from this rank's
            avg_rows = self.M_global // (torch.distributed.get_world_size(self._group) if self._dist else 1)
            lib = _L.lib()
            native_tc = lib.drsa_step_workspace_bytes(max(avg_rows, 1), d, m, num_concepts, _L.PREC_TC_F16) >= 0
            if precision == "auto" and native_tc:
                precision = _auto_precision(self.M_global, avg_rows, d, m, num_concepts)
            plan, pad_prec = None, precision
            if not native_tc and m == d and (precision in ("tc", "tc_split", "tc_hilo", "tc_dc", "tc32") or
                                             (precision == "auto" and avg_rows >= 65536)):
                plan = _pad_plan(d, m, num_concepts, avg_rows)
                if plan is not None and precision == "auto":
                    pad_prec = _auto_precision(self.M_global, avg_rows, plan[0], plan[1], num_concepts)
            elif not native_tc and m == d and d > 64 and d % 32 != 0 and precision in ("auto", "fp32"):
                # exact fp32 arithmetic, padded only so that the fused finish kernel (d, m multiples of 32) applies:
                # the un-fused retraction costs ~30 launches (0.5 ms at d = 100)
                plan, pad_prec = _pad_plan(d, m, num_concepts, avg_rows), "fp32"
            if plan is None and precision == "auto":
                precision = "fp32"
            if plan is not None:
                dp, mp, dkp = plan
                cols = (torch.arange(m, device=self.device) // self.d_k) * dkp + torch.arange(m, device=self.device) % self.d_k
                Up = torch.zeros(dp, mp, device=self.device)
                Up[:d, cols] = self._Uw
                free = torch.ones(mp, dtype=torch.bool, device=self.device)
                free[cols] = False
                extra = torch.nonzero(free).flatten()          # m' - m padding columns: unit vectors in the new channels
                Up[d + torch.arange(extra.numel(), device=self.device), extra] = 1.0
                self._pad = dict(d=d, m=m, dp=dp, mp=mp, cols=cols)
                self._Uw = Up
                act_p = torch.zeros(rows, dp, device=self.device)
                act_p[:, :d] = _f32c(self._act_src, self.device)
                ctx_p = torch.zeros(rows, dp, device=self.device)
                ctx_p[:, :d] = _f32c(self._ctx_src, self.device)
                self._rows = _RowPass(act_p, ctx_p, dp, mp, num_concepts, pad_prec, self.device)
                del act_p, ctx_p
                d, m = dp, mp
            else:
                self._rows = _RowPass(self._act_src, self._ctx_src, d, m, num_concepts, precision, self.device)
                if not self._rows.is_tc:
                    self._act_dev, self._ctx_dev = self._rows.A, self._rows.C
        self.precision = self._rows.precision
        self.retraction_iters, self.retraction_tol = retraction_iters, retraction_tol
        self._px = None
        self.exchange = "none"
        if self._dist:
            self.exchange = self._setup_exchange(exchange, d, m, num_concepts)
        self.use_cuda_graph = use_cuda_graph and (not self._dist or self._px is not None)
        self.obj_history: Optional[np.ndarray] = None
        self.last_status: Optional[np.ndarray] = None

    @property
    def act_vecs(self) -> torch.Tensor:
        """The activation rows on the device in fp32 (drsa.py:70); materialised on first access."""
        if self._act_dev is None:
            self._act_dev = _f32c(self._act_src, self.device)
        return self._act_dev

    @property
    def ctx_vecs(self) -> torch.Tensor:
        """The context rows on the device in fp32 (drsa.py:71); materialised on first access."""
        if self._ctx_dev is None:
            self._ctx_dev = _f32c(self._ctx_src, self.device)
        return self._ctx_dev

    @property
    def U(self) -> torch.Tensor:
        """The d x m projection matrix (drsa.py:66).  When the problem runs zero-padded on the tensor cores this is a copy
        of the original block of the working matrix."""
        if self._pad is None:
            return self._Uw
        return self._Uw[: self._pad["d"]][:, self._pad["cols"]].contiguous()

    @U.setter
    def U(self, value: torch.Tensor) -> None:
        v = _f32c(value, self.device)
        if self._pad is None:
            self._Uw = v.clone()
        else:
            self._Uw[: self._pad["d"]][:, self._pad["cols"]] = v
        self._graph = None

    def _setup_exchange(self, exchange: str, d: int, m: int, K: int) -> str:
        dist = torch.distributed
        if exchange not in ("auto", "p2p", "p2p_symm", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p', 'p2p_symm' or 'nccl'")
        exchange = os.environ.get("DRSA_EXCHANGE", exchange)
        world = dist.get_world_size(self._group)
        if exchange == "nccl" or world < 2:
            return "nccl"
        nbytes = int(_L.lib().drsa_exchange_bytes(d, m, K, world))
        on_nccl = dist.get_backend(self._group) == "nccl"
        if nbytes < 0 or not on_nccl:
            if exchange != "auto":
                raise _L.DRSAError(f"exchange='{exchange}' needs an NCCL group of at most {_L.MAX_PEERS} ranks and d, m "
                                   "multiples of 32")
            return "nccl"
        with torch.cuda.device(self.device):
            self._px = _PeerExchange.get(self._group, nbytes, self.device, "symm" if exchange == "p2p_symm" else "ipc")
        if self._px is None:
            if exchange != "auto":
                raise _L.DRSAError(f"exchange='{exchange}': the peer buffers could not be set up on every rank")
            return "nccl"
        return "p2p_symm" if exchange == "p2p_symm" else "p2p"

    # ------------------------------------------------------------------ one step
    def _step(self, obj_log: torch.Tensor, log_index: int, update: bool, correct: bool = True) -> None:
        self._rows.step(self._Uw, correct)
        if self.retraction == "qr":
            r = self._rows
            _L.check(r.lib.drsa_finish_step_qr(_ptr(r.sums), self.M_global, _ptr(self._Uw), r.d, r.m, r.K,
                                               _ptr(self._Uw) if update else None, _ptr(obj_log), log_index, _ptr(r.status),
                                               _ptr(r.ws_fin), r.ws_fin.numel(), _stream()), "drsa_finish_step_qr")
            return
        if self._dist and self._px is None:
            torch.distributed.all_reduce(self._rows.sums, group=self._group)   # d*m + K floats over NVLink
        self._rows.finish(self._Uw, self.M_global, obj_log, log_index, update, self.retraction_iters,
                          self.retraction_tol, self._px)

    def run(self, steps: int = 2000, save: bool = True) -> None:
        """``steps`` ascent steps; the objective is logged before every update plus once at the end
        (drsa.py:84-117), then U and the statistics are written (drsa.py:119-120)."""
        with torch.cuda.device(self.device):
            self._rows.split_u(self._Uw)
            self.reset_log(steps + 1)
            self.enqueue_steps(steps)
            self._step(self._obj_log, -1, False, False)   # final evaluation, no update (drsa.py:109-117)
            hist = self._obj_log[: steps + 1].cpu().numpy()   # the only device->host copy of the loop
            self.last_status = self._rows.status.cpu().numpy()
        self.obj_history = hist
        if int(self.last_status[1]) != 0:
            import warnings
            warnings.warn(f"DRSA: the polar retraction did not converge within {self.retraction_iters} sweeps in "
                          f"{int(self.last_status[1])} step(s); raise retraction_iters")
        if save and self.path_to_model is not None:
            self.save_model()
            self.save_train_stats([np.asarray(v) for v in hist])

    def reset_log(self, capacity: int) -> None:
        """(Re)allocate the device-side objective log and rewind its cursor."""
        log = getattr(self, "_obj_log", None)
        if log is None or log.numel() < capacity:
            self._obj_log = torch.zeros(max(capacity, 4096), dtype=torch.float32, device=self.device)
            self._graph = None                            # the log pointer is baked into the graph
        self._rows.status.zero_()

    def enqueue_steps(self, n: int) -> None:
        """Enqueue ``n`` ascent steps on the current stream without any host synchronisation.  Each
        step appends its objective to the device log.  One step is captured in a CUDA graph once and replayed
        ('tc_dc': two graphs, the step that re-evaluates the correction and the plain one)."""
        if n <= 0:
            return
        every = dc_every(self.M_global) if self._rows.dc else 1
        count = getattr(self, "_steps_done", 0)
        if not (self.use_cuda_graph and n >= 4):
            for _ in range(n):
                self._step(self._obj_log, -1, True, count % every == 0)
                count += 1
            self._steps_done = count
            return
        if getattr(self, "_graph", None) is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._step(self._obj_log, -1, True, count % every == 0)   # warm-up outside the graph (lazy module init)
            torch.cuda.current_stream().wait_stream(side)
            count += 1
            n -= 1
            # raw capture_begin / capture_end on the side stream: the torch.cuda.graph() context manager runs gc.collect()
            # and torch.cuda.empty_cache() first, which hands GBs of cached blocks back to the driver (100 ms .. 1 s of
            # cudaFree / cudaMalloc inside the end-to-end time); a step allocates nothing, so no private pool is needed
            graphs = {}
            with torch.cuda.stream(side):
                for correct in ((True, False) if self._rows.dc else (True,)):
                    graph = torch.cuda.CUDAGraph()
                    graph.capture_begin()
                    try:
                        self._step(self._obj_log, -1, True, correct)      # capture does not execute
                    finally:
                        graph.capture_end()
                    graphs[correct] = graph
            torch.cuda.current_stream().wait_stream(side)
            self._graph = graphs
        for _ in range(n):
            self._graph[count % every == 0].replay()
            count += 1
        self._steps_done = count

    # ------------------------------------------------------------------ objective (static, differentiable)
    @staticmethod
    def obj_val(act_vecs: torch.Tensor, context_vecs: torch.Tensor, U: torch.Tensor, obj_fn=None,
                num_concepts: int = 4, d_k: Optional[int] = None) -> torch.Tensor:
        """DRSA objective of drsa.py:123-155 as a 0-dim tensor on U's device, differentiable w.r.t. U
        (the backward is the fused gradient of the same row pass).  ``obj_fn`` is accepted for
        signature compatibility; the pooling of drsa.py:224-238 is what the kernel implements."""
        if obj_fn is not None and obj_fn is not objective_fn:
            raise _L.DRSAError("obj_val implements the reference's objective_fn only")
        if d_k is not None and d_k * num_concepts != U.size(1):
            raise ValueError("num_concepts * d_k must equal the number of columns of U")
        return _ObjVal.apply(act_vecs, context_vecs, U, num_concepts)

    # ------------------------------------------------------------------ outputs (formats of drsa.py:157-168)
    def save_train_stats(self, obj_arr: List[np.ndarray]) -> None:
        import pandas as pd
        pd.DataFrame({"loss": obj_arr}).to_csv(os.path.join(self.path_to_model, "train_stats.csv"))

    def save_model(self) -> None:
        with open(os.path.join(self.path_to_model, "projection_matrix.pkl"), "wb") as file:
            pickle.dump(self.U.detach().cpu().numpy(), file)


class _ObjVal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act, cvec, U, K):
        dev = _as_device(U.device if U.is_cuda else _DEFAULT_DEVICE)
        with torch.cuda.device(dev):
            A, Cv, Uc = _f32c(act, dev), _f32c(cvec, dev), _f32c(U, dev)
            rows = _RowPass(A, Cv, Uc.size(0), Uc.size(1), K, "fp32", dev)
            rows.step(Uc)
            obj = torch.zeros(1, dtype=torch.float32, device=dev)
            rows.finish(Uc, A.size(0), obj, 0, False, 1, 1e-6)
        ctx.save_for_backward(rows.sums, obj)
        ctx.shape = (Uc.size(0), Uc.size(1), K, A.size(0))
        return obj[0].to(device=U.device, dtype=U.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        sums, obj = ctx.saved_tensors
        d, m, K, M = ctx.shape
        X = sums[: d * m].view(d, m)
        q = torch.sqrt(sums[d * m:] / M)
        coef = torch.sqrt(obj) / (K * M * q ** 1.5)
        grad = X * coef.repeat_interleave(m // K)[None, :]
        return None, None, (grad * grad_out.to(grad.device)).to(device=grad_out.device, dtype=grad_out.dtype), None


# ---------------------------------------------------------------------- module-level functions
def generalized_fmean(x: torch.Tensor, p: float = 0.5) -> torch.Tensor:
    """F-mean with F(t) = t^p over dim 0 (drsa.py:171-182)."""
    return torch.pow(torch.mean(torch.pow(x, p), dim=0), 1 / p)


@torch.no_grad()
def project_grad(gradient: torch.Tensor, U: torch.Tensor) -> torch.Tensor:
    """Unused in the reference as well (drsa.py:186-198); kept for API completeness."""
    return gradient - torch.matmul(torch.matmul(U.T, gradient), U.T)


@torch.no_grad()
def orthogonalize(U: torch.Tensor, max_iters: int = 24, tol: float = 1e-6, method: str = "polar") -> torch.Tensor:
    """U (U^T U)^(-1/2) -- the reference's retraction (drsa.py:201-221), computed on the device.  ``method='qr'`` returns Q of
    the thin QR factorisation with diag(R) > 0 instead (comparison only: not the reference's retraction)."""
    dev = _as_device(U.device if U.is_cuda else _DEFAULT_DEVICE)
    lib = _L.lib()
    with torch.cuda.device(dev):
        Y = _f32c(U, dev)
        d, m = Y.shape
        out = torch.empty_like(Y)
        ws = torch.empty(int(_L.check(lib.drsa_finish_workspace_bytes(d, m))), dtype=torch.uint8, device=dev)
        status = torch.zeros(4, dtype=torch.int32, device=dev)
        if method == "qr":
            _L.check(lib.drsa_qr_retract(_ptr(Y), d, m, _ptr(out), _ptr(status), _ptr(ws), ws.numel(), _stream()),
                     "drsa_qr_retract")
            return out.to(U.dtype)
        _L.check(lib.drsa_polar_retract(_ptr(Y), d, m, _ptr(out), max_iters, tol, _ptr(status), _ptr(ws), ws.numel(),
                                        _stream()), "drsa_polar_retract")
    return out.to(U.dtype)


def objective_fn(input: torch.Tensor) -> torch.Tensor:
    """Soft-max pooling over rows (p = 2), soft-min pooling over concepts (p = 0.5) -- drsa.py:224-238."""
    return generalized_fmean(generalized_fmean(input, 2), 0.5)


def main(activation_vecs: torch.Tensor, context_vecs: torch.Tensor, model_root: str, num_concepts: int = 4,
         steps: int = 2000, runs: int = 3, seed: int = 42, device=_DEFAULT_DEVICE, **optimizer_kwargs) -> None:
    """Several runs from differently permuted random orthogonal starts (drsa.py:241-300): numpy is
    seeded once, ``ortho_group.rvs(d)`` draws U, each run permutes the columns of the *previous*
    run's start (the permutations compound) and writes into ``model_root/run{r}``."""
    from scipy.stats import ortho_group

    np.random.seed(seed)
    device = _as_device(device)
    print(f"Starting DRSA training on device {device} ...")
    d = activation_vecs.size(-1)
    U = ortho_group.rvs(d)
    print("Orthogonal projection matrix U of size (%2d x %2d)" % (d, d))
    for run in range(1, runs + 1):
        model_path = os.path.join(model_root, f"run{run}")
        os.makedirs(model_path, exist_ok=True)
        mask = np.random.permutation(d)
        U = U[:, mask]
        print("-" * 20, f"\nStarting RUN {run}")
        opt = SubspaceOptimizer(torch.tensor(U, dtype=activation_vecs.dtype), activation_vecs, context_vecs,
                                model_path, num_concepts=num_concepts, device=device, **optimizer_kwargs)
        opt.run(steps=steps)
    print("Done!")
