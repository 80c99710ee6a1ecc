#!/usr/bin/env bash
# cfg-2 weak scaling on ONE 8-GPU box, N = 1, 2, 4, 8 back to back (short end-to-end leg: the lines are for `value`).
set -u
mkdir -p gpurun_out
ARGS="--steps 200 --warmup 5 --no-lrp --no-modes --no-parity --no-cpu-baseline --e2e-steps 200"
timeout 150 python bench.py --gpus 1 $ARGS > gpurun_out/r02_scale_cfg2_n1.log 2>&1; echo "n1 exit $?"
for N in 2 4 8; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29850 + N)) \
    bench.py --gpus $N $ARGS > gpurun_out/r02_scale_cfg2_n$N.log 2>&1; echo "n$N exit $?"
done
for N in 1 2 4 8; do grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": [0-9]*' gpurun_out/r02_scale_cfg2_n$N.log | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_scale_cfg2_n$N.log | head -1; grep -o '"row_pass_per_rank": \[[^]]*\]' gpurun_out/r02_scale_cfg2_n$N.log; done
