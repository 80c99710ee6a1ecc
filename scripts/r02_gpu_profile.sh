#!/usr/bin/env bash
# Round-2 single-GPU measurement session: cfg 3 / cfg 4 / cfg 1 / cfg 5 bench lines, then the ncu launch list and full captures
# of the same command (each only after the plain command exited 0).
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg3 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg3.log 2>&1; echo "cfg3 exit $?"
timeout 1500 python bench.py --workload cfg4 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_cfg4.log 2>&1; echo "cfg4 exit $?"
timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 5 > gpurun_out/r02_bench_cfg1.log 2>&1; echo "cfg1 exit $?"
timeout 900 python bench.py --workload cfg5 > gpurun_out/r02_bench_cfg5_n1.log 2>&1; echo "cfg5 exit $?"
PROF="python bench.py --steps 17 --warmup 3 --no-cpu-baseline --e2e-steps 4 --no-graph --no-parity --no-modes --lrp-samples 64"
$PROF > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench_nograph.csv $PROF > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list exit $?"
$PROF > gpurun_out/r02_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"drsa_tc_step_kernel" -s 4 -c 4 -o gpurun_out/r02_prof_tc $PROF > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:"finish_fused_kernel" -s 4 -c 2 -o gpurun_out/r02_prof_finish $PROF > gpurun_out/r02_ncu_full_finish.log 2>&1
echo "finish capture exit $?"
