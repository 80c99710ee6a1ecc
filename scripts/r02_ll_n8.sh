#!/usr/bin/env bash
# 8-GPU check of the flag-in-data peer exchange: phase profile of the finish kernel and the cfg-2 weak-scaling line.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29841 scripts/p2p_exchange_profile.py > gpurun_out/r02_ll_exchange_profile_n$N.log 2>&1; echo "profile exit $?"; grep "^rank 0" gpurun_out/r02_ll_exchange_profile_n$N.log
timeout 300 $TR --master-port 29842 bench.py --gpus $N --steps 200 --warmup 5 --no-lrp --no-modes > gpurun_out/r02_ll_bench_cfg2_n$N.log 2>&1; echo "bench exit $?"; tail -c 600 gpurun_out/r02_ll_bench_cfg2_n$N.log | head -c 300; grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": [0-9]*' gpurun_out/r02_ll_bench_cfg2_n$N.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_ll_bench_cfg2_n$N.log
