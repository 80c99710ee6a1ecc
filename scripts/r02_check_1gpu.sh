#!/usr/bin/env bash
# Single-GPU re-validation after a kernel change: full GPU suite, smoke, finish-kernel phase stamps, the driver's bench command.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02b_pytest_gpu.log; cat gpurun_out/r02b_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -2 gpurun_out/r02b_smoke.log
python scripts/finish_phase_profile.py 256 > gpurun_out/r02b_finish_phase_stamps.log 2>&1; cat gpurun_out/r02b_finish_phase_stamps.log | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_cfg2.log 2>&1; echo "bench exit $?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_cfg2_reference_arm.log 2>&1; echo "reference arm exit $?"
python bench.py --steps 200 --warmup 5 --no-lrp --no-modes --no-parity > gpurun_out/r02b_bench_cfg2_k200.log 2>&1; echo "bench K=200 exit $?"
