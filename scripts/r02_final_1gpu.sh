#!/usr/bin/env bash
# Final single-GPU session of round 2: full GPU test suite, smoke, the driver's bench command and its reference arm, then the
# ncu launch list and full captures of the same command (each only after the plain command exited 0).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2.log 2>&1; echo "bench exit $?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2_reference_arm.log 2>&1; echo "reference arm exit $?"
python scripts/finish_phase_profile.py 256 > gpurun_out/r02_finish_phase_stamps.log 2>&1
PROF="python bench.py --steps 33 --warmup 3 --no-cpu-baseline --e2e-steps 4 --no-graph --no-parity --no-modes --lrp-samples 64"
$PROF > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_nograph.csv $PROF > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list exit $?"
$PROF > gpurun_out/r02_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"drsa_tc_step_kernel" -s 4 -c 4 -o gpurun_out/r02_prof_tc $PROF > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:"finish_fused_kernel" -s 4 -c 2 -o gpurun_out/r02_prof_finish $PROF > gpurun_out/r02_ncu_full_finish.log 2>&1
echo "finish capture exit $?"
