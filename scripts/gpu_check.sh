#!/usr/bin/env bash
# One GPU-box session: parity tests, smoke, bench, then (NCU=1) ncu launch list + one full capture of the top kernel.
set -u
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${T:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n "${TAIL:-8}" gpurun_out/$name.log; }
run drsa python -m pytest tests/test_gpu_drsa.py -q -m gpu -s
run lrp python -m pytest tests/test_gpu_lrp.py tests/test_gpu_logmel.py -q -m gpu -s
run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=900 run bench python bench.py --steps 200 --warmup 5 --e2e-steps 500
if [ "${NCU:-0}" = "1" ]; then
PROF="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 4 --no-graph --lrp-samples 64"
echo "=== ncu"
$PROF > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$PROF > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"drsa_tc_step_kernel|conv3x3_tc_kernel" -s 4 -c 4 -o gpurun_out/prof_tc $PROF > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc_kernel" -s 1 -c 2 -o gpurun_out/prof_conv $PROF > gpurun_out/ncu_full_conv.log 2>&1
echo "conv capture exit $?"
fi
