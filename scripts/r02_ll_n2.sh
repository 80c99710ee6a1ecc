#!/usr/bin/env bash
# 2-GPU check of the flag-in-data peer exchange: distributed tests, phase profile of the finish kernel, a short cfg-2 bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r02_ll_pytest_n2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02_ll_pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29831 scripts/p2p_exchange_profile.py > gpurun_out/r02_ll_exchange_profile_n2.log 2>&1; echo "profile exit $?"; grep "^rank" gpurun_out/r02_ll_exchange_profile_n2.log
timeout 600 $TR --master-port 29832 bench.py --gpus 2 --steps 200 --warmup 5 --no-lrp --no-modes > gpurun_out/r02_ll_bench_cfg2_n2.log 2>&1; echo "bench exit $?"; tail -c 1500 gpurun_out/r02_ll_bench_cfg2_n2.log
