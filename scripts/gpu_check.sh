#!/usr/bin/env bash
# One GPU-box session: selftests, parity tests (each group in its own process), smoke, short bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n "${TAIL:-15}" gpurun_out/$name.log; }
run selftest python -m pytest tests/test_gpu_drsa.py -q -m gpu -x -k "selftest" -s
run fp32 python -m pytest tests/test_gpu_drsa.py -q -m gpu -k "fp32 or polar or obj_val or subspace or (golden and fp32)" -s
run tc python -m pytest tests/test_gpu_drsa.py -q -m gpu -k "tensor_core or (golden and tc) or large" -s
run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=900 run bench python bench.py --steps 20 --warmup 3 --e2e-steps 200
