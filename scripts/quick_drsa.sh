#!/usr/bin/env bash
# quick GPU check of the DRSA row pass: parity tests, short bench line, per-phase cycle counters
python -m pytest tests/test_gpu_drsa.py -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 200 --warmup 5 --no-lrp --no-cpu-baseline --e2e-steps 50 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['step_breakdown_ms'], d['clocks'])"
python scripts/tc_phase_profile.py 2>&1 | tail -6
python scripts/tc_phase_profile.py tc_split 2>&1 | tail -6
