#!/usr/bin/env bash
# Round-2 multi-GPU session (N = number of visible GPUs): cfg 2 weak scaling, cfg 3 strong scaling, cfg 5 pipeline, cfg 4.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29821 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2_n$N.log 2>&1; echo "cfg2 exit $?"
timeout 600 $TR --master-port 29822 bench.py --gpus $N --workload cfg3 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg3_n$N.log 2>&1; echo "cfg3 exit $?"
timeout 900 $TR --master-port 29823 bench.py --gpus $N --workload cfg5 > gpurun_out/r02_bench_cfg5_n$N.log 2>&1; echo "cfg5 exit $?"
timeout 900 $TR --master-port 29824 bench.py --gpus $N --workload cfg4 --steps 10 --warmup 5 > gpurun_out/r02_bench_cfg4_n$N.log 2>&1; echo "cfg4 exit $?"
timeout 600 $TR --master-port 29825 bench.py --gpus $N --workload cfg1 --steps 200 --warmup 5 --no-lrp > gpurun_out/r02_bench_cfg1_n$N.log 2>&1; echo "cfg1 exit $?"
for f in cfg2 cfg3 cfg5 cfg4 cfg1; do tail -c 400 gpurun_out/r02_bench_${f}_n$N.log; echo; done
