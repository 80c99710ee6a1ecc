set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_lrp.py -q -m gpu -x -s > gpurun_out/lrp.log 2>&1; echo "lrp exit $?"; tail -15 gpurun_out/lrp.log
python scripts/debug_lrp_tc.py > gpurun_out/dbg_lrp.log 2>&1; echo "dbg exit $?"; tail -50 gpurun_out/dbg_lrp.log
