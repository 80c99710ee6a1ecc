#!/usr/bin/env bash
# compute-sanitizer over the hand-written pipelines on small shapes; logs -> gpurun_out/sanitize_*.log (copied to profiles/).
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for what in drsa lrp; do
  timeout 1500 $CS --tool memcheck --print-limit 20 --log-file gpurun_out/sanitize_memcheck_$what.log python scripts/sanitize_target.py $what > gpurun_out/sanitize_memcheck_$what.out 2>&1
  echo "memcheck $what exit $?"; tail -3 gpurun_out/sanitize_memcheck_$what.log
done
for what in drsa lrp; do
  timeout 1500 $CS --tool racecheck --print-limit 20 --log-file gpurun_out/sanitize_racecheck_$what.log python scripts/sanitize_target.py $what > gpurun_out/sanitize_racecheck_$what.out 2>&1
  echo "racecheck $what exit $?"; tail -3 gpurun_out/sanitize_racecheck_$what.log
done
timeout 900 $CS --tool synccheck --print-limit 20 --log-file gpurun_out/sanitize_synccheck_drsa.log python scripts/sanitize_target.py drsa > gpurun_out/sanitize_synccheck_drsa.out 2>&1
echo "synccheck drsa exit $?"; tail -3 gpurun_out/sanitize_synccheck_drsa.log
