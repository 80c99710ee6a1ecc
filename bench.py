#!/usr/bin/env python
"""Benchmark of the DRSA/LRP hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Workload (config.workload = "cfg2"): BASELINE.json configs[1] -- DRSA at the last conv of the
widened genre CNN: N = 10 000 samples x P = 64 positions -> M = 640 000 (activation, context)
rows per GPU, d = 256, K = 4 concepts of d_k = 64.  A "step" is one DRSA optimisation step
(row pass + d*m all-reduce + ascent + polar retraction).  Rows shard across ranks with the
per-GPU row count fixed (weak scaling); `value` is whole-job throughput in cfg2-sized steps per
second, i.e. (rows processed per step by all ranks / 640 000) * steps/s, which equals plain
steps/s at N = 1.
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
METRIC = "DRSA steps/s (cfg2: M=640k rows x d=256, K=4; LRP context vecs/s reported in 'lrp')"
UNIT = "steps/s"


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


def synth_rows_cuda(M: int, d: int, seed: int, device):
    """SURVEY 8(d) synthetic pairs, generated on the device: A = relu(randn)*mask, C = randn*(A>0),
    each normalised like normalize_vectors."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    A = torch.relu(torch.randn(M, d, generator=g, device=device)) * (torch.rand(M, d, generator=g, device=device) < 0.7)
    C = torch.randn(M, d, generator=g, device=device) * (A > 0)
    nv = lambda v: v / torch.sqrt(torch.mean(v * v)) / d ** 0.25
    return nv(A).contiguous(), nv(C).contiguous()


def synth_U0(d: int, seed: int = 5):
    """Random orthogonal start (Q of a seeded Gaussian matrix), generated on the CPU and shared by both arms."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.linalg.qr(torch.randn(d, d, generator=g))[0].contiguous()


# =========================================================================== this repo's arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cxai.xai.drsa.drsa import SubspaceOptimizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M, d, K = CFG2["samples"] * CFG2["positions"], CFG2["d"], CFG2["K"]
    if args.rows:
        M = args.rows
    if args.d:
        d = args.d
    if args.K:
        K = args.K
    custom = bool(args.rows or args.d or args.K)
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
    opt._rows.split_u(opt._Uw)
    opt.reset_log(args.warmup + args.steps + 8)
    warm = max(args.warmup, 4 if not args.no_graph else 3)      # >= 4 so that the CUDA graph of one step is captured here,
    opt.enqueue_steps(warm)                                     # not inside the timed region
    barrier()
    if rank == 0:
        t_wait = time.time() + 5.0
        while not sampler.rows and time.time() < t_wait:
            time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    opt.enqueue_steps(args.steps)                   # EXACTLY K steps
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    # ---- dominant kernel alone (row pass = tcgen05 kernel + partial reduce), CUDA events on its stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        opt._rows.step(opt._Uw)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        opt._rows.step(opt._Uw)
    k1.record()
    torch.cuda.synchronize()
    ms_kernel = k0.elapsed_time(k1) / args.steps
    # ---- the replicated tail of a step alone (ascent + polar retraction)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Utmp = opt._Uw.clone()
    f0.record()
    for _ in range(args.steps):
        opt._rows.finish(Utmp, opt.M_global, None, 0, True, opt.retraction_iters, opt.retraction_tol)
    f1.record()
    torch.cuda.synchronize()
    ms_finish = f0.elapsed_time(f1) / args.steps
    status = opt._rows.status.cpu().numpy().tolist()
    per_rank = None
    if world > 1:          # the step is synchronous: the slowest rank's row pass sets the pace (power capping differs per GPU)
        tk = torch.tensor([ms_kernel], dtype=torch.float64, device=dev)
        allk = [torch.zeros_like(tk) for _ in range(world)]
        dist.all_gather(allk, tk)
        per_rank = [round(float(v.item()), 4) for v in allk]
    lrp = None if args.no_lrp else lrp_throughput(args, dev, rank, world, barrier)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API with HOST buffers: construct (H2D + pack) + run + D2H
    e2e_steps = args.e2e_steps
    Ah, Ch = A.cpu().pin_memory(), C.cpu().pin_memory()
    del opt
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

    if rank == 0:
        ms_per_step = ms_total / args.steps
        scale = world if custom else (M * world) / float(CFG2["samples"] * CFG2["positions"])
        value = scale * 1000.0 / ms_per_step
        flops = 8.0 * M * d * m
        achieved = flops / (ms_kernel * 1e-3) / 1e12
        is_tc = opt2.precision in ("tc", "tc_split")
        mma_factor = {"tc": 1.0, "tc_split": 1.5}.get(opt2.precision, 1.0)
        elem = 2 if is_tc else 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"tc": "f16 operands / f32 accumulate", "tc_split": "f16 operands (U split hi+lo) / f32 accumulate"}.get(opt2.precision, "f32"),
            "data": "synthetic",
            "config": {"workload": "cfg2" if not custom else f"custom rows={M} d={d} K={K} (steps/s of THIS shape per GPU x GPUs)",
                       "rows_per_gpu": M, "d": d, "m": m, "K": K, "d_k": m // K, "precision": opt2.precision, "cuda_graph": bool(opt2.use_cuda_graph),
                       "exchange": {"none": "single rank", "nccl": "NCCL all-reduce of d*m+K floats per step",
                                    "p2p": "all-reduce fused into the finish kernel over NVLink peer memory (cudaIpc buffers)",
                                    "p2p_symm": "all-reduce fused into the finish kernel over NVLink peer memory (torch symmetric memory)"}[opt2.exchange],
                       "l2_policy": f"inputs larger than L2 ({2 * M * d * elem / 1e6:.0f} MB of rows per step vs 126 MB L2)",
                       "retraction_sweeps_last_step": status[0], "retraction_not_converged": status[1]},
            "rows_per_s": M * world * 1000.0 / ms_per_step,
            "gpu_launches": int(args.steps * launches_per_step(opt2)),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tf_sustained"], "traffic": TRAFFIC_BYTES_PER_LAUNCH if (is_tc and not custom) else None,
                         "kernel": "drsa_tc_step_kernel (+ tc_reduce_kernel)" if is_tc else "sgemm_kernel chain",
                         "kernel_ms": ms_kernel, "algorithmic_flop_per_launch": flops,
                         "algorithmic_bytes_per_launch": 2.0 * M * d * elem,
                         "executed_mma_flop_per_launch": flops * mma_factor * (1.5 if (is_tc and d > 256) else 1.0),
                         "hbm_frac": (2.0 * M * d * elem / (ms_kernel * 1e-3) / 1e9) / peaks["hbm"],
                         "peak_source": peaks["source"] + "; sustained bf16 figure (kernel timed in a loop)",
                         "share_of_step": ms_kernel / ms_per_step},
            "step_breakdown_ms": {"row_pass": ms_kernel, "ascent_and_retraction": ms_finish,
                                  **({"row_pass_per_rank": per_rank} if per_rank else {})},
            "lrp": lrp,
            "e2e": {"value": scale * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_steps,
                    "d2h_bytes_per_step": d2h / e2e_steps, "steps": e2e_steps,
                    "what": "SubspaceOptimizer(U0, A_host_pinned, C_host_pinned).run(steps) + U.cpu(): H2D of all rows, "
                            "fp16 pack, steps, final objective, D2H of U and the objective history, wall clock"},
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
    + normalise, through the public cxai API (get_intermediate + gather_context_pairs)."""
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

    def once(xb):
        a, R = pp.get_intermediate(net, xb, comp, layer, 0)
        return pp.gather_context_pairs(a, R, None, normalize=True)

    once(x)                                     # warm-up with the full batch (allocator + plan cache)
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
    """The reference's stage 1 on the host cores: get_intermediate over the oracle's restatement of zennit's rules (general
    multi-pass Gamma, backward continued to the input as zennit does), fp32, cfg-2 CNN, n samples."""
    import torch
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
TRAFFIC_BYTES_PER_LAUNCH = 662.8e6      # 656.3 MB read + ~6.5 MB written (profiles/r01_ncu_full_drsa_tc_step_kernel_v4.csv)


def launches_per_step(opt) -> int:
    """Kernels of libdrsa_b200.so launched per DRSA step (counted from the host code in csrc/)."""
    row = 2 if opt.precision in ("tc", "tc_split") else 8 * max(1, -(-opt.act_vecs.size(0) // (1 << 18)))
    return row + 1                      # row pass (+ partial reduce) and the fused cooperative finish kernel


# =========================================================================== CPU baseline / reference arm
def cpu_step_time(M_sample: int, d: int, K: int, budget_s: float, threads: int):
    """Times the reference's algorithm (torch CPU ops in the reference's order: obj_val, autograd
    backward, orthogonalize with the fp64 eigh -- oracle/drsa_ref.step_autograd restates drsa.py:84-104)."""
    import torch
    from oracle import drsa_ref
    torch.set_num_threads(threads)
    A, C = drsa_ref.synth_pairs(M_sample, d, 20262, structured=False)
    U = drsa_ref.synth_U0(d, seed=5)
    _, _, U = drsa_ref.step_autograd(A, C, U, K)          # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        _, _, U = drsa_ref.step_autograd(A, C, U, K)
        times.append(time.perf_counter() - t0)
    return sum(times) / len(times), len(times)


def cpu_baseline(M: int, d: int, K: int, budget_s: float = 15.0):
    import torch
    threads = os.cpu_count() or 1
    M_sample = min(M, 64_000)
    t, n = cpu_step_time(M_sample, d, K, budget_s, threads)
    return {"value": (M_sample / M) / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} steps of the reference algorithm (oracle/drsa_ref.step_autograd, torch {torch.__version__} CPU, "
                      f"{threads} threads) on {M_sample} of the {M} rows; steps/s scaled linearly in rows",
            "sample_ms_per_step": t * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    M, d, K = CFG2["samples"] * CFG2["positions"], CFG2["d"], CFG2["K"]
    if args.rows:
        M = args.rows
    if args.d:
        d = args.d
    if args.K:
        K = args.K
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    M_sample = min(M, 64_000)
    from oracle import drsa_ref
    torch.set_num_threads(threads)
    A, C = drsa_ref.synth_pairs(M_sample, d, 20262, structured=False)
    U = drsa_ref.synth_U0(d, seed=5)
    for _ in range(max(1, min(args.warmup, 3))):
        _, _, U = drsa_ref.step_autograd(A, C, U, K)
    steps = min(args.steps, 40)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, U = drsa_ref.step_autograd(A, C, U, K)
    t = (time.perf_counter() - t0) / steps
    value = (M_sample / M) / t
    sample = (f"{steps} steps of the reference algorithm (drsa.py:84-104 restated in oracle/drsa_ref.step_autograd; "
              f"/root/reference is a Python repo that cannot travel to the GPU box) on {M_sample} of {M} rows, torch CPU, "
              f"{threads} threads; steps/s scaled linearly in rows")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(1, min(args.warmup, 3)), "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2" if not args.rows else f"custom rows={M}", "rows_per_gpu": M, "d": d, "m": d, "K": K,
                   "d_k": d // K},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tc", choices=["tc", "tc_split", "fp32", "auto"])
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (default: cfg2 = 640000)")
    ap.add_argument("--d", type=int, default=0, help="override the split-layer width (default: cfg2 = 256); cfg4 uses 512")
    ap.add_argument("--K", type=int, default=0, help="override the number of concepts (default: cfg2 = 4); cfg4 uses 8")
    ap.add_argument("--e2e-steps", type=int, default=2000,
                    help="steps of the end-to-end call (default: the reference's run(steps=2000), drsa.py:76)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lrp", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "p2p_symm", "nccl"],
                    help="how the row sums are joined across ranks (see SubspaceOptimizer)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly (used for the ncu captures)")
    ap.add_argument("--lrp-samples", type=int, default=256)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
