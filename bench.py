#!/usr/bin/env python
"""Benchmark of the DRSA/LRP hot path on B200 (contract: see the task statement / DESIGN.md section 7).

    python bench.py --gpus N --steps K --warmup W [--workload cfg1|cfg2|cfg3|cfg4|cfg5]   # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W                        # the reference's CPU code

Workloads = BASELINE.json configs.  A "step" is one DRSA optimisation step (row pass + exchange of d*m+K
floats + ascent + polar retraction) over ALL rows of the workload.
  cfg1  toy scale, the reference's CPU-runnable case: M = 16 000 rows (N = 1k x P = 16), d = 64, K = 4 per GPU (weak)
  cfg2  (default; the configuration the metric is quoted on) genre CNN, last conv: N = 10 000 samples x P = 64 positions
        -> M = 640 000 rows per GPU, d = 256, K = 4 x 64 (weak scaling: rows per GPU fixed; `value` counts cfg2-sized
        steps per second of the whole job = (rows of all ranks / 640 000) * steps/s, plain steps/s at N = 1)
  cfg3  N = 100 000 samples: M = 6.4 M rows in total, sharded over the ranks (strong scaling), d = 256, K = 4
  cfg4  wide split layer: d = 512, K = 8 x 64, M = 12.8 M rows in total (51.2 M on 8 GPUs), strong scaling
  cfg5  end to end per class: for 10 classes, 10 000 synthetic log-mel spectrograms -> CNN forward -> LRP to features[33]
        -> (a, c) pairs -> normalise -> DRSA (K = 4, --e2e-steps steps); value = DRSA steps per second of wall clock
        through the whole pipeline
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG2 = dict(samples=10_000, positions=64, d=256, K=4)
WORKLOADS = {
    "cfg1": dict(rows_per_gpu=16_000, d=64, K=4, scaling="weak", unit_rows=16_000),
    "cfg2": dict(rows_per_gpu=640_000, d=256, K=4, scaling="weak", unit_rows=640_000),
    "cfg3": dict(rows_total=6_400_000, d=256, K=4, scaling="strong"),
    "cfg4": dict(rows_total=12_800_000, rows_total_8=51_200_000, d=512, K=8, scaling="strong"),
}
METRIC = "DRSA steps/s (cfg2: M=640k rows x d=256, K=4; LRP context vecs/s reported in 'lrp')"
UNIT = "steps/s"
OBJ_TOL, ANGLE_TOL = 1e-4, 1e-3        # BASELINE.json north_star


def workload_shape(args, world: int):
    """(rows on this rank, d, K, scaling, value scale = cfg-sized steps per step of the whole job, label)."""
    if args.rows or args.d or args.K:
        M, d, K = args.rows or 640_000, args.d or 256, args.K or 4
        return M, d, K, "weak", float(world), f"custom rows={M} d={d} K={K} (steps/s of THIS shape per GPU x GPUs)"
    w = WORKLOADS[args.workload]
    if w["scaling"] == "weak":
        return w["rows_per_gpu"], w["d"], w["K"], "weak", float(world), args.workload
    total = w.get("rows_total_8", w["rows_total"]) if world >= 8 else w["rows_total"]
    return total // world, w["d"], w["K"], "strong", 1.0, args.workload


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi sampling DURING the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_rows_cuda(M: int, d: int, seed: int, device, chunk: int = 1 << 20):
    """SURVEY 8(d) synthetic pairs, generated on the device in row chunks (cfg 4 is 26 GB per matrix): A = relu(randn)*mask,
    C = randn*(A>0), each normalised like normalize_vectors (global statistic)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    A = torch.empty(M, d, device=device)
    C = torch.empty(M, d, device=device)
    ssA = torch.zeros((), dtype=torch.float64, device=device)
    ssC = torch.zeros((), dtype=torch.float64, device=device)
    for r0 in range(0, M, chunk):
        r1 = min(M, r0 + chunk)
        a = torch.relu(torch.randn(r1 - r0, d, generator=g, device=device))
        a *= (torch.rand(r1 - r0, d, generator=g, device=device) < 0.7)
        c = torch.randn(r1 - r0, d, generator=g, device=device)
        c *= (a > 0)
        A[r0:r1], C[r0:r1] = a, c
        ssA += (a.double() ** 2).sum()
        ssC += (c.double() ** 2).sum()
    A *= float(1.0 / (torch.sqrt(ssA / (M * d)) * d ** 0.25))
    C *= float(1.0 / (torch.sqrt(ssC / (M * d)) * d ** 0.25))
    return A, C


def synth_U0(d: int, seed: int = 5):
    """Random orthogonal start (Q of a seeded Gaussian matrix), generated on the CPU and shared by both arms."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.linalg.qr(torch.randn(d, d, generator=g))[0].contiguous()


def principal_angle(U1, U2, K: int) -> float:
    """Largest principal angle (rad) between matching concept subspaces: asin of the singular values of (I - Q1 Q1^T) Q2 in
    fp64 on the host (acos of the cosines has a 1e-3 rad noise floor in fp32)."""
    import torch
    U1, U2 = U1.detach().double().cpu(), U2.detach().double().cpu()
    d, m = U1.shape
    dk, worst = m // K, 0.0
    eye = torch.eye(d, dtype=torch.float64)
    for k in range(K):
        Q1 = torch.linalg.qr(U1[:, k * dk:(k + 1) * dk])[0]
        Q2 = torch.linalg.qr(U2[:, k * dk:(k + 1) * dk])[0]
        sv = torch.linalg.svdvals((eye - Q1 @ Q1.T) @ Q2)
        worst = max(worst, float(torch.asin(sv.clamp(max=1.0)).max()))
    return worst


def launches_in(opt, first_step: int, n: int) -> int:
    """Kernels of libdrsa_b200.so launched by steps first_step .. first_step+n-1 (counted from the host code in csrc/):
    row pass = tcgen05 kernel + partial reduce (fp32: sgemm chain), 'tc_dc' adds the combine kernel and, on a correction
    step, a second (hi + lo) row pass; the finish step is one cooperative kernel."""
    from cxai.xai.drsa import drsa as D
    prec = opt.precision
    if prec == "fp32":
        row = 1 if (opt._rows.d <= 64 and opt._rows.m <= 64 and opt._rows.K <= 4) else 8 * max(1, -(-opt._rows.M // (1 << 18)))
        return n * (row + 1)
    if prec != "tc_dc":
        return n * 3
    every = D.dc_every(opt.M_global)
    corr = sum(1 for i in range(first_step, first_step + n) if i % every == 0)
    return corr * 6 + (n - corr) * 4


def time_steps(opt, n: int, warm: int = 6):
    """Device time of n steps (ms per step), CUDA events on the launching stream, after `warm` steps (graph capture)."""
    import torch
    opt.reset_log(warm + n + 8)
    opt.enqueue_steps(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opt.enqueue_steps(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def parity_block(A, C, U0, K, precision, dev, steps, timed_objs, timed_U, budget_s=90.0):
    """The timed arithmetic mode against the on-device fp32 CUDA-core path (itself pinned to the reference's golden
    trajectories over 2 000 steps, tests/test_gpu_drsa_long.py) on the SAME rows and start, over the e2e horizon: largest
    relative objective error over all steps and the final principal angle.  If the fp32 run of the full shape does not fit
    the time budget, both runs are repeated on a row subset (stated)."""
    import numpy as np
    import torch
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    M = A.size(0)
    probe = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device=dev, precision="fp32")
    probe._rows.split_u(probe._Uw)
    ms32 = time_steps(probe, 3, warm=2)
    del probe
    rows = M
    if ms32 * 1e-3 * steps > budget_s:
        rows = max(65536, int(M * budget_s / (ms32 * 1e-3 * steps)) // 65536 * 65536)
    if rows < M:
        A, C = A[:rows].contiguous(), C[:rows].contiguous()
        o = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device=dev, precision=precision)
        o.run(steps=steps, save=False)
        timed_objs, timed_U = o.obj_history, o.U.clone()
        del o
    t0 = time.perf_counter()
    ref = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device=dev, precision="fp32")
    ref.run(steps=steps, save=False)
    t32 = time.perf_counter() - t0
    rel = float(np.max(np.abs(timed_objs - ref.obj_history) / np.abs(ref.obj_history)))
    ang = principal_angle(timed_U, ref.U, K)
    return {"mode": precision, "rows": rows, "steps": steps,
            "against": "on-device fp32 CUDA-core path (pinned to the reference's 2000-step golden trajectories in "
                       "tests/test_gpu_drsa_long.py), same rows and start matrix",
            "max_rel_objective_err": rel, "principal_angle_rad": ang, "tol": {"objective": OBJ_TOL, "angle_rad": ANGLE_TOL},
            "ok": bool(rel < OBJ_TOL and ang < ANGLE_TOL), "fp32_path_steps_per_s": steps / t32,
            **({"note": f"fp32 run of the full {M} rows exceeds the {budget_s:.0f} s budget: both modes run on the first {rows} rows"}
               if rows < M else {})}


# =========================================================================== this repo's arm
def correction_window(start: int, steps: int, every: int):
    """(untimed steps to run first, correction steps then inside the window): the smallest shift of a window of `steps`
    steps starting at step index `start` after which it contains round(steps / every) indices that are multiples of `every`
    (the steps on which 'tc_dc' re-evaluates its correction).  every = 0: the mode has no such steps."""
    if every <= 0:
        return 0, 0
    inside = lambda s0: sum(1 for i in range(s0, s0 + steps) if i % every == 0)      # noqa: E731
    target = int(round(steps / every))
    align = next(a for a in range(every) if inside(start + a) == target)
    return align, target


def workload_config(label, M, world, d, m, K):
    """The `config` block of the JSON line: the workload only, identical in both arms."""
    fp16_mb, fp32_mb = 2 * M * d * 2 / 1e6, 2 * M * d * 4 / 1e6
    return {"workload": label, "rows_per_gpu": M, "rows_total": M * world, "d": d, "m": m, "K": K, "d_k": m // K,
            "l2_policy": (f"inputs larger than L2 ({fp32_mb:.0f} MB of fp32 rows per GPU, {fp16_mb:.0f} MB as packed fp16 planes, "
                          f"read once per step, vs 126 MB L2)") if fp16_mb > 126
            else "rows fit in L2 (cfg 1 is the reference's toy scale: latency-bound)"}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cxai.xai.drsa.drsa import SubspaceOptimizer
    from cxai.xai.drsa import drsa as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "cfg5" and not (args.rows or args.d or args.K):
        return run_cfg5(args, dev, rank, world)
    M, d, K, scaling, scale_units, label = workload_shape(args, world)
    m = d
    A, C = synth_rows_cuda(M, d, 20262 + rank, dev)
    U0 = synth_U0(d, seed=5)
    peaks = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                             # started early: nvidia-smi start-up must not land in the timed region
    opt = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device=dev, precision=args.precision,
                            use_cuda_graph=not args.no_graph, exchange=args.exchange)
    precision = opt.precision
    opt._rows.split_u(opt._Uw)
    warm = max(args.warmup, 4 if not args.no_graph else 3)      # >= 4 so that the CUDA graph(s) of a step are captured here,
    opt.reset_log(warm + args.steps + 8 + 64)                   # not inside the timed region (+ alignment steps, below)
    opt.enqueue_steps(warm)
    # 'tc_dc' re-evaluates its correction (one extra hi + lo row pass) every `every` steps.  The timed window must carry its
    # share of those steps -- the integer nearest to K / every -- wherever the K steps happen to start: a few more untimed
    # steps move the window accordingly (K = 20, every = 32: one correction step inside, 0.625 expected).
    every = D.dc_every(opt.M_global) if opt._rows.dc else 0
    align, corrections = correction_window(opt._steps_done, args.steps, every)
    opt.enqueue_steps(align)
    barrier()
    if rank == 0:
        t_wait = time.time() + 5.0
        while not sampler.rows and time.time() < t_wait:
            time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    first_timed = opt._steps_done
    e0.record()
    opt.enqueue_steps(args.steps)                   # EXACTLY K steps
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    n_launch = launches_in(opt, first_timed, args.steps)
    # ---- dominant kernel alone: the single-plane row pass (tcgen05 kernel + partial reduce), CUDA events on its stream
    is_tc = precision != "fp32"
    kcode = {"tc_dc": D._L.PREC_TC_F16}.get(precision, opt._rows.prec_code)
    kout = opt._rows.sums_raw if opt._rows.dc else opt._rows.sums
    krun = (lambda: opt._rows._row_sums(opt._Uw, kcode, kout)) if is_tc else (lambda: opt._rows.step(opt._Uw))
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        krun()
    torch.cuda.synchronize()
    nk = max(args.steps, 20)
    k0.record()
    for _ in range(nk):
        krun()
    k1.record()
    torch.cuda.synchronize()
    ms_kernel = k0.elapsed_time(k1) / nk
    ms_hilo = None
    if opt._rows.dc:                                # the hi + lo row pass of a correction step
        k0.record()
        for _ in range(10):
            opt._rows._row_sums(opt._Uw, D._L.PREC_TC_F16_AC2, opt._rows.sums)
        k1.record()
        torch.cuda.synchronize()
        ms_hilo = k0.elapsed_time(k1) / 10
    # ---- the replicated tail of a step alone (ascent + polar retraction)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Utmp = opt._Uw.clone()
    f0.record()
    for _ in range(nk):
        opt._rows.finish(Utmp, opt.M_global, None, 0, True, opt.retraction_iters, opt.retraction_tol)
    f1.record()
    torch.cuda.synchronize()
    ms_finish = f0.elapsed_time(f1) / nk
    status = opt._rows.status.cpu().numpy().tolist()
    per_rank = None
    if world > 1:          # the step is synchronous: the slowest rank's row pass sets the pace (power capping differs per GPU)
        tk = torch.tensor([ms_kernel], dtype=torch.float64, device=dev)
        allk = [torch.zeros_like(tk) for _ in range(world)]
        dist.all_gather(allk, tk)
        per_rank = [round(float(v.item()), 4) for v in allk]
    exchange = opt.exchange
    use_graph = bool(opt.use_cuda_graph)
    del opt, Utmp
    lrp = None if (args.no_lrp or args.workload in ("cfg3", "cfg4")) else lrp_throughput(args, dev, rank, world, barrier)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the other arithmetic modes at the same shape (device-timed, 1 GPU)
    modes = None
    if world == 1 and not args.no_modes:
        modes = {}
        for pm in ("tc", "tc_dc", "tc_hilo", "tc32", "fp32"):
            if pm == "fp32" and 8.0 * M * d * m > 4e12:          # the CUDA-core path would take seconds per step
                continue
            try:
                o = SubspaceOptimizer(U0, A, C, None, num_concepts=K, device=dev, precision=pm)
            except Exception:                       # noqa: BLE001 -- mode not available for this shape (tc32 at d = 512)
                continue
            o._rows.split_u(o._Uw)
            msm = time_steps(o, 16 if pm == "fp32" else 64)
            modes[pm] = {"steps_per_s": 1000.0 / msm, "ms_per_step": msm}
            del o
        modes["note"] = ("'tc' = rows once in fp16; 'tc_dc' = 'tc' + deferred correction of the row rounding every "
                         f"{D.dc_every(M * world)} steps; 'tc_hilo' = rows as hi + lo fp16 planes; 'tc32' = rows and U hi + lo "
                         "(fp32-class operands); 'fp32' = CUDA cores.  'auto' never picks 'tc'.")

    # ---- cfg 2 with the reference's own sampling of L = 20 positions per spectrogram (getdrsadata.py:131): M = 200 000 rows
    sampled = None
    if world == 1 and args.workload == "cfg2" and label == "cfg2" and not args.no_modes:
        Ms = M * 20 // 64
        o = SubspaceOptimizer(U0, A[:Ms], C[:Ms], None, num_concepts=K, device=dev, precision=args.precision)
        o._rows.split_u(o._Uw)
        mss = time_steps(o, 64)
        sampled = {"rows": Ms, "positions_per_sample": 20, "precision": o.precision, "steps_per_s": 1000.0 / mss,
                   "ms_per_step": mss, "what": "the first 200 000 of the 640 000 rows (rows are i.i.d. in the synthetic workload)"}
        del o

    # ---- end to end through the public API with HOST buffers: construct (H2D + pack) + run + D2H
    e2e_steps = args.e2e_steps
    # host copies of the rows; workloads whose rows exceed 16 GB (cfg 4: 52 GB) use the first rows only -- the host side of
    # the box is not sized for a pinned copy of everything -- and the steps/s are scaled by the row fraction (stated)
    e2e_rows = M if 2.0 * M * d * 4 <= 16e9 else int(16e9 / (2.0 * d * 4)) // 65536 * 65536
    if e2e_rows < M and e2e_steps > 400:
        e2e_steps = 400
    Ah, Ch = A[:e2e_rows].cpu().pin_memory(), C[:e2e_rows].cpu().pin_memory()
    torch.cuda.synchronize()
    # one untimed call first (like the warm-up steps above): the stage-1 benchmark left the caching allocator fragmented
    # and the first construction after it pays for cudaFree/cudaMalloc round trips that are not part of the path
    optw = SubspaceOptimizer(U0, Ah, Ch, None, num_concepts=K, device=dev, precision=args.precision,
                             use_cuda_graph=not args.no_graph, exchange=args.exchange)
    optw.run(steps=8, save=False)
    del optw
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    opt2 = SubspaceOptimizer(U0, Ah, Ch, None, num_concepts=K, device=dev, precision=args.precision,
                             use_cuda_graph=not args.no_graph, exchange=args.exchange)
    opt2.run(steps=e2e_steps, save=False)
    U_host = opt2.U.cpu()
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    objs = opt2.obj_history
    h2d = (Ah.numel() + Ch.numel() + U0.numel()) * 4
    d2h = U_host.numel() * 4 + (e2e_steps + 1) * 4
    U_dev = opt2.U.clone()
    del opt2, Ah, Ch
    replicas = None
    if world > 1:          # every rank ran the same deterministic finish kernel on the same sums: U must be BIT-identical
        digest = torch.stack([U_dev.view(torch.int32).to(torch.int64).sum(), U_dev.double().abs().sum().view(torch.int64)])
        alld = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(alld, digest)
        replicas = {"ranks": world, "bit_identical": bool(all(torch.equal(v, alld[0]) for v in alld))}
    parity = None
    if world == 1 and not args.no_parity and precision != "fp32":
        if e2e_rows < M:
            A, C = A[:e2e_rows].contiguous(), C[:e2e_rows].contiguous()       # the rows the end-to-end run used
        parity = parity_block(A, C, U0, K, precision, dev, e2e_steps, objs, U_dev, budget_s=args.parity_budget)

    if rank == 0:
        ms_per_step = ms_total / args.steps
        unit_rows = WORKLOADS.get(args.workload, {}).get("unit_rows")
        scale = scale_units if (scaling == "strong" or unit_rows is None or label.startswith("custom")) else \
            (M * world) / float(unit_rows)
        value = scale * 1000.0 / ms_per_step
        flops = 8.0 * M * d * m
        achieved = flops / (ms_kernel * 1e-3) / 1e12
        elem = 2 if is_tc else 4
        dtype = {"fp32": "f32", "tc": "f16 operands / f32 accumulate",
                 "tc_split": "f16 operands (U hi+lo) / f32 accumulate",
                 "tc_hilo": "f16 hi+lo row planes (22 bit) / f32 accumulate",
                 "tc_dc": "f16 operands / f32 accumulate + deferred 22-bit correction of the row rounding",
                 "tc32": "f16 hi+lo operands (22 bit rows and U) / f32 accumulate"}[precision]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            # `config` names the workload and is the same dict in the reference arm; what THIS arm ran with is in `run`
            "config": workload_config(label, M, world, d, m, K),
            "run": {"precision": precision, "precision_requested": args.precision, "cuda_graph": use_graph,
                    "exchange": {"none": "single rank", "nccl": "NCCL all-reduce of d*m+K floats per step",
                                 "p2p": "all-reduce fused into the finish kernel over NVLink peer memory (cudaIpc buffers)",
                                 "p2p_symm": "all-reduce fused into the finish kernel over NVLink peer memory (torch symmetric memory)"}[exchange],
                    "row_bytes_read_per_step": 2 * M * d * elem,
                    "correction_every": every or None, "correction_steps_in_timed_region": corrections,
                    "untimed_alignment_steps_after_warmup": align,
                    "retraction_sweeps_last_step": status[0], "retraction_not_converged": status[1]},
            "rows_per_s": M * world * 1000.0 / ms_per_step,
            "gpu_launches": int(n_launch),
            "roofline": {"bound": "tensor" if d >= 128 else "hbm", "achieved": achieved, "peak": peaks["tf_burst"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
                         "frac_of_sustained_peak": achieved / peaks["tf_sustained"],
                         "traffic": TRAFFIC_BYTES_PER_LAUNCH if (is_tc and args.workload == "cfg2" and not label.startswith("custom")) else None,
                         "kernel": "drsa_tc_step_kernel, single-plane rows (+ tc_reduce_kernel)" if is_tc else "fp32 row-pass kernels",
                         "kernel_ms": ms_kernel, "algorithmic_flop_per_launch": flops,
                         "algorithmic_bytes_per_launch": 2.0 * M * d * elem,
                         "executed_mma_flop_per_launch": flops * {"tc_split": 1.25, "tc_hilo": 2.0, "tc32": 2.5}.get(precision, 1.0)
                         * (1.5 if (is_tc and d > 256) else 1.0),
                         "hbm_frac": (2.0 * M * d * elem / (ms_kernel * 1e-3) / 1e9) / peaks["hbm"],
                         "peak_source": peaks["source"] + f"; burst bf16 figure (kernel timed alone, {nk} back-to-back launches); "
                                        f"sustained figure {peaks['tf_sustained']} TFLOP/s in frac_of_sustained_peak",
                         "share_of_step": ms_kernel / ms_per_step},
            "step_breakdown_ms": {"row_pass": ms_kernel, "ascent_and_retraction": ms_finish,
                                  **({"row_pass_hi_lo_on_correction_steps": ms_hilo, "correction_every": D.dc_every(M * world)} if ms_hilo else {}),
                                  **({"row_pass_per_rank": per_rank} if per_rank else {})},
            "lrp": lrp,
            "e2e": {"value": scale * e2e_steps / t_e2e * (e2e_rows / float(M)), "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_steps,
                    "d2h_bytes_per_step": d2h / e2e_steps, "steps": e2e_steps,
                    **({"rows": e2e_rows, "note": f"run on the first {e2e_rows} of {M} rows, steps/s scaled by the row fraction"}
                       if e2e_rows < M else {}),
                    "what": "SubspaceOptimizer(U0, A_host_pinned, C_host_pinned).run(steps) + U.cpu(): H2D of all rows, "
                            "fp16 pack, steps, final objective, D2H of U and the objective history, wall clock"},
            "parity": parity, "modes": modes, "cfg2_L20": sampled, "replicas": replicas,
            "clocks": clocks,
            "objective_first_last": [float(objs[0]), float(objs[-1])],
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(M, d, K, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def build_cfg2_model(device):
    """BASELINE cfg 2 CNN: arch A (getdrsadata.py:72-73) with the last block widened to 256 filters, BatchNorm,
    input 128x256, PyTorch default init under torch.manual_seed(0), eval mode."""
    import torch
    from cxai.model.create_model import VGGType
    torch.manual_seed(0)
    net = VGGType(n_filters=[64, 64, 100, 128, 256], pool_kernels=[(2, 4), (2, 2), (2, 2), (2, 2), (2, 2)], n_dense=100,
                  n_classes=10, dropout=0.3, block_depth=2, dense_depth=2, input_size=(128, 256), conv_bn=True,
                  dense_bn=True)
    return net.eval().to(device)


def lrp_throughput(args, dev, rank, world, barrier):
    """LRP context vectors / s: log-mel batch -> CNN forward -> LRP down to features[33] -> gather + c=R/(a+1e-7)
    + normalise, through the public cxai API (preprocessing.extract_context_pairs)."""
    import torch
    import torch.distributed as dist
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.drsa import preprocessing as pp
    net = build_cfg2_model(dev)
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    n = args.lrp_samples
    g = torch.Generator(device=dev).manual_seed(20262 + rank)
    x = (1.2 * torch.randn(n, 1, 128, 256, generator=g, device=dev) - 1.5).clamp(min=-4.0)
    layer = net.features[33]

    def once(xb):          # get_intermediate + gather + c = R/(a+1e-7) + normalise, rows written straight from the NHWC planes
        return pp.extract_context_pairs(net, xb, comp, 33, 0, normalize=True, device=dev)

    for _ in range(3):                          # warm-up with the full batch: plan, allocator, and the CUDA graph of an engine
        once(x)                                 # pass (first call plain launches, second call capture, then replays)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    act, ctx = once(x)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # end to end from pinned host memory, result (vectors) read back
    xh = x.cpu().pin_memory()
    act, ctx = once(xh.to(dev, non_blocking=True))          # warm the pinned-copy path
    barrier()
    t0 = time.perf_counter()
    act, ctx = once(xh.to(dev, non_blocking=True))
    _ = act[:1].cpu()
    barrier()
    t_e2e = time.perf_counter() - t0
    # full-depth relevance maps (compute_relevances, attribute.py:70-108: the other output of the same pass)
    from cxai.xai.explain.attribute import compute_relevances
    nf = min(n, 128)
    for _ in range(3):                          # plain launches, graph capture, replay
        compute_relevances(net, x[:nf], comp, class_idx=0)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    Rin = compute_relevances(net, x[:nf], comp, class_idx=0)
    f1.record()
    barrier()
    ms_full = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_full, op=dist.ReduceOp.MAX)
    ms_full = float(ms_full.item())
    del Rin
    P = act.shape[0] // n
    flops = 2 * 1775.5e6 * n           # 2 x MACs of the widened arch A forward per sample
    out = {"metric": "LRP context vecs/s", "value": n * P * world / (ms * 1e-3), "unit": "vectors/s",
           "samples_per_gpu": n, "positions": P, "d": int(act.shape[1]), "ms": ms,
           "e2e_value": n * P * world / t_e2e, "h2d_bytes": int(xh.numel() * 4),
           "forward_tflops": flops / (ms * 1e-3) / 1e12,
           "kernel": "conv3x3_tc_kernel (tcgen05 implicit GEMM, TMA im2col, fp16 hi/lo operands, max-pool fused into the "
                     "epilogue below the split layer); first conv, dense head and pool routing on CUDA cores",
           "samples_per_engine_pass": min(n, 256),
           "relevance_maps": {"metric": "LRP relevance maps at the input (compute_relevances, full-depth backward)",
                              "value": nf * world / (ms_full * 1e-3), "unit": "maps/s", "samples_per_gpu": nf, "ms": ms_full,
                              "algorithmic_tflops": 3 * 2 * 1775.5e6 * nf / (ms_full * 1e-3) / 1e12}}
    if rank == 0:
        out["roofline"] = lrp_conv_roofline(dev, ms / max(1, n / 64.0))
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = lrp_cpu_baseline(P)
    return out


def lrp_conv_roofline(dev, stage_ms_per_minibatch):
    """The dominant kernel of stage 1 alone: conv3x3_tc_kernel<2, resident weights, halo boxes> on the 64 -> 64 layer at
    128 x 256 (68 % of the forward MACs) with the (2, 4) max-pool of create_model.py:120 fused into its epilogue, as the
    engine launches it below the split layer; 64 samples per launch, CUDA events on its stream."""
    import torch
    from drsa_audio_b200 import _lib as L
    lib = L.lib()
    B, H, W, Cc = 64, 128, 256, 64
    g = torch.Generator(device=dev).manual_seed(1)
    xh = torch.rand(B, H, W, Cc, generator=g, device=dev).half()
    xl = (torch.rand(B, H, W, Cc, generator=g, device=dev) * 1e-4).half()
    wt = torch.randn(9, Cc, Cc, generator=g, device=dev) / 24.0
    wh = torch.empty(9, Cc, Cc, dtype=torch.float16, device=dev); wl = torch.empty_like(wh)
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.lrp_tc_split_f16(wt.data_ptr(), wt.numel(), wh.data_ptr(), wl.data_ptr(), s))
    bias = torch.zeros(Cc, device=dev)
    yh = torch.empty(B, H // 2, W // 4, Cc, dtype=torch.float16, device=dev); yl = torch.empty_like(yh)
    err = torch.zeros(1, dtype=torch.int32, device=dev)

    def run():
        L.check(lib.lrp_tc_conv3x3_forward_pool(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), bias.data_ptr(), B, H,
                                                W, Cc, Cc, Cc, 1, 2, 4, yh.data_ptr(), yl.data_ptr(), None, err.data_ptr(), s))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / 10
    flops = 2.0 * 9 * Cc * Cc * H * W * B
    peaks = _peaks()
    ach = flops / (kms * 1e-3) / 1e12
    in_bytes, out_bytes = 2.0 * B * H * W * Cc * 2, 2.0 * B * (H // 2) * (W // 4) * Cc * 2
    return {"bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
            "traffic": LRP_CONV_TRAFFIC_BYTES_PER_LAUNCH,
            "kernel": "conv3x3_tc_kernel<2, resident weights, halo boxes> + fused (2,4) max-pool (64->64 @128x256, 64 samples)",
            "kernel_ms": kms, "algorithmic_flop_per_launch": flops, "executed_mma_flop_per_launch": 3 * flops,
            "algorithmic_bytes_per_launch": in_bytes + out_bytes,
            "hbm_frac": ((in_bytes + out_bytes) / (kms * 1e-3) / 1e9) / peaks["hbm"],
            "share_of_stage": kms / stage_ms_per_minibatch,
            "note": "fp32-class accuracy needs three fp16 products per MAC (hi*hi + lo*hi + hi*lo, SURVEY H4), so the "
                    "algorithmic fraction is bounded by 1/3; executed MMA flop / peak = 3 x frac",
            "l2_policy": "inputs larger than L2 (537 MB of activation planes per launch)",
            "peak_source": peaks["source"] + "; sustained bf16 figure"}


# dram__bytes_read.sum + dram__bytes_write.sum of that launch (profiles/r01_ncu_full_conv_first_fusedpool_v6.csv):
# 537.1 MB read + 61.3 MB written vs 536.9 + 67.1 MB algorithmic
LRP_CONV_TRAFFIC_BYTES_PER_LAUNCH = 598.4e6


def lrp_cpu_baseline(P: int, n: int = 4):
    """The reference's stage 1 on the host cores.  With ``oracle/_ref`` present (byte copy of the reference sources made by
    oracle/make_ref.py) the reference's OWN get_intermediate runs in a subprocess on the mini-zennit restatement
    (kind = "reference"); otherwise the oracle port (oracle/lrp_ref.py, general multi-pass rules, kind = "port")."""
    import torch
    if os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "cxai", "xai", "drsa", "preprocessing.py")):
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_baseline.py"), "lrp", str(n)],
                                 capture_output=True, text=True, timeout=600, cwd=ROOT)
            res = json.loads(out.stdout.strip().splitlines()[-1])
            if "value" in res:
                return res
        except Exception:                           # noqa: BLE001 -- fall through to the port
            pass
    from oracle import lrp_ref
    from cxai.utils.constants import lrp_name_map_6s
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    net = lrp_ref.genre_model(seed=0, last=256, input_size=(128, 256))
    x = lrp_ref.synth_logmel(n, 128, 256, 20262)
    nm = lrp_name_map_6s()
    lrp_ref.get_intermediate(net, x[:1], nm, net.features[33], 0, dtype=torch.float32)       # warm-up
    t0 = time.perf_counter()
    lrp_ref.get_intermediate(net, x, nm, net.features[33], 0, dtype=torch.float32)
    t = time.perf_counter() - t0
    return {"value": n * P / t, "unit": "vectors/s", "cores": threads, "kind": "port",
            "sample": f"{n} samples of 128x256 through oracle/lrp_ref.get_intermediate (torch {torch.__version__} CPU, fp32, "
                      f"{threads} threads), {P} positions each", "sample_s": t}


# dram__bytes_read.sum + dram__bytes_write.sum of one drsa_tc_step_kernel launch at cfg2 from the committed
# ncu --set full capture (profiles/); None until a capture exists.
TRAFFIC_BYTES_PER_LAUNCH = 661.4e6      # 656.3 MB read + ~5.1 MB written (profiles/r02_ncu_full_drsa_tc_step_kernel.csv)


# =========================================================================== CPU baseline / reference arm
def reference_stepper(threads: int):
    """(step function (A, C, U, K) -> U_next, kind, description).  With ``oracle/_ref`` present this is the reference's
    own code, unmodified: ``SubspaceOptimizer.obj_val`` (drsa.py:123-155) + ``backward`` (:100) + ``orthogonalize``
    (:201-221) exactly as ``run`` (:84-104) strings them together; otherwise the oracle port of the same lines."""
    import torch
    torch.set_num_threads(threads)
    from oracle import make_ref
    if make_ref.available():
        ref = make_ref.load_drsa()

        def step(A, C, U, K):
            U = U.detach().requires_grad_(True)                                   # drsa.py:86-88
            obj = ref.SubspaceOptimizer.obj_val(A, C, U, ref.objective_fn, K, U.size(1) // K)   # :91-98
            obj.backward()                                                        # :100
            with torch.no_grad():
                return ref.orthogonalize(U + U.grad)                              # :102
        return step, "reference", "the reference's own drsa.py (byte copy in oracle/_ref): obj_val + backward + orthogonalize"
    from oracle import drsa_ref
    return (lambda A, C, U, K: drsa_ref.step_autograd(A, C, U, K)[2]), "port", \
        "the reference algorithm restated in oracle/drsa_ref.step_autograd (oracle/_ref absent)"


def synth_rows_cpu(M: int, d: int, seed: int):
    import torch
    g = torch.Generator().manual_seed(seed)
    A = torch.relu(torch.randn(M, d, generator=g)) * (torch.rand(M, d, generator=g) < 0.7)
    C = torch.randn(M, d, generator=g) * (A > 0)
    nv = lambda v: v / torch.sqrt(torch.mean(v * v)) / d ** 0.25
    return nv(A).contiguous(), nv(C).contiguous()


def cpu_baseline(M: int, d: int, K: int, budget_s: float = 15.0):
    """The reference's step on the host cores on a BOUNDED sample of the workload: at most 640 000 rows (cfg 2 in full) and
    about `budget_s` seconds; larger workloads are scaled linearly in the row count (the step is O(M d^2) + O(d^3))."""
    import torch
    threads = os.cpu_count() or 1
    step, kind, what = reference_stepper(threads)
    M_sample = min(M, 640_000)
    A, C = synth_rows_cpu(M_sample, d, 20262)
    U = synth_U0(d, seed=5)
    U = step(A, C, U, K)                            # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        U = step(A, C, U, K)
        times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return {"value": (M_sample / M) / t, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{len(times)} steps of {what}, torch {torch.__version__} CPU, {threads} threads, on {M_sample} of the {M} rows"
                      + ("" if M_sample == M else "; steps/s scaled linearly in rows"),
            "sample_ms_per_step": t * 1e3}


def run_reference(args):
    """The reference's CPU implementation of the path on the box's host cores, all threads, on the workload's FULL row count
    where that is at most cfg 2's (larger workloads: 640 000 rows, scaled linearly, stated); --steps / --warmup honoured."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "cfg5" and not (args.rows or args.d or args.K):
        args.workload = "cfg2"                      # the DRSA stage of one class of cfg 5 is the cfg 2 step
    M, d, K, scaling, scale_units, label = workload_shape(args, world)
    threads = os.cpu_count() or 1
    step, kind, what = reference_stepper(threads)
    M_sample = min(M, 640_000)
    A, C = synth_rows_cpu(M_sample, d, 20262)
    U = synth_U0(d, seed=5)
    for _ in range(max(1, args.warmup)):
        U = step(A, C, U, K)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        U = step(A, C, U, K)
    t = (time.perf_counter() - t0) / args.steps
    value = (M_sample / M) / t
    if scaling == "weak":
        value *= 1.0                                # one host, one problem of the per-GPU size: the N = 1 figure
    sample = (f"{args.steps} steps of {what} on {M_sample} of {M} rows, torch {torch.__version__} CPU, {threads} threads"
              + ("" if M_sample == M else "; steps/s scaled linearly in rows"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(label, M, world, d, d, K),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# =========================================================================== cfg 5: the per-class pipeline end to end
def run_cfg5(args, dev, rank, world):
    """10 classes x 10 000 synthetic spectrograms -> forward -> LRP to features[33] with class_idx = c -> pairs ->
    normalise -> DRSA (K = 4, --e2e-steps steps per class).  --steps / --warmup do not apply (one pass = 10 classes)."""
    import torch
    import torch.distributed as dist
    from cxai.utils.constants import lrp_name_map_6s
    from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
    from cxai.xai.drsa.cluster.optsubspaces import all_classes_pipeline, class_pipeline
    n_class, steps, classes = args.cfg5_samples, args.e2e_steps, 10
    net = build_cfg2_model(dev)
    comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
    n_local = n_class // world
    g = torch.Generator(device=dev).manual_seed(20265 + rank)
    batch = lambda: (1.2 * torch.randn(n_local, 1, 128, 256, generator=g, device=dev) - 1.5).clamp(min=-4.0)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    # warm-up: the same schedule on 64 samples per class and rank, 8 steps -- plans, engine-pass graphs, the NCCL channels of
    # the all-to-all and of the sub-groups and their peer-memory exchanges are set up here, not inside the timed region
    # (the first all-to-all of a process group alone took 8.5 s of the 11.4 s a cold 8-GPU run needed)
    warm = {c: batch()[:64].contiguous() for c in range(classes)}
    for _ in range(2):
        all_classes_pipeline(net, warm, comp, 33, None, num_concepts=4, steps=8, precision=args.precision)
    del warm
    data = {c: batch() for c in range(classes)}                 # this rank's spectrograms of every class, resident
    sync()
    t0 = time.perf_counter()
    timings = {} if args.cfg5_timings else None
    out = all_classes_pipeline(net, data, comp, 33, None, num_concepts=4, steps=steps, precision=args.precision,
                               **({"timings": timings} if world > 1 else {}))
    sync()
    wall = time.perf_counter() - t0
    tw = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall = float(tw.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        vecs = classes * n_local * world * 64
        objs = [(float(out[c][1][0]), float(out[c][1][-1])) for c in range(classes)]
        print(json.dumps({
            "metric": METRIC, "value": classes * steps / wall, "unit": UNIT, "n_gpus": world, "steps": classes * steps,
            "warmup": 8, "ms_per_step": wall * 1e3 / (classes * steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16 hi+lo operands / f32 accumulate (stage 1), DRSA precision 'auto'", "data": "synthetic",
            "config": {"workload": "cfg5", "classes": classes, "samples_per_class": n_local * world, "positions": 64,
                       "d": 256, "K": 4, "drsa_steps_per_class": steps, "precision": args.precision,
                       "l2_policy": "inputs larger than L2 (1.3 GB of spectrograms and 1.3 GB of rows per class)"},
            "wall_s": wall, "phases_s_rank0": timings, "context_vectors": vecs, "vectors_per_s_whole_pipeline": vecs / wall,
            "gpu_launches": None, "clocks": clocks, "objective_first_last_per_class": objs,
            "e2e": {"value": classes * steps / wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                    "what": "spectrograms resident on the device (synthetic); wall clock over all 10 classes incl. the D2H "
                            "of every objective history"},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--precision", default="auto", choices=["auto", "tc", "tc_dc", "tc_hilo", "tc32", "tc_split", "fp32"],
                    help="arithmetic of the row pass; 'auto' (what a user gets) = 'tc_dc' at cfg 2")
    ap.add_argument("--rows", type=int, default=0, help="custom shape: rows per GPU")
    ap.add_argument("--d", type=int, default=0, help="custom shape: split-layer width")
    ap.add_argument("--K", type=int, default=0, help="custom shape: number of concepts")
    ap.add_argument("--e2e-steps", type=int, default=2000,
                    help="steps of the end-to-end call and of the parity check (default: the reference's run(steps=2000), drsa.py:76)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--parity-budget", type=float, default=90.0, help="seconds the fp32 yardstick run of the parity block may take")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lrp", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-modes", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "p2p_symm", "nccl"],
                    help="how the row sums are joined across ranks (see SubspaceOptimizer)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly (used for the ncu captures)")
    ap.add_argument("--lrp-samples", type=int, default=256)
    ap.add_argument("--cfg5-samples", type=int, default=10_000, help="cfg 5: samples per class (over all ranks)")
    ap.add_argument("--cfg5-timings", action="store_true", help="cfg 5: wall-clock breakdown of rank 0 (adds synchronisations)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
