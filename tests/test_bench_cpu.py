"""Host logic of bench.py that needs no GPU: workload table, the `config` block shared by both arms, the placement of the
timed window relative to 'tc_dc' correction steps, and the reference arm (`--impl reference`) end to end on the smallest
workload."""
import json
import os
import subprocess
import sys
from types import SimpleNamespace

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench          # noqa: E402


def _args(**kw):
    base = dict(workload="cfg2", rows=0, d=0, K=0)
    base.update(kw)
    return SimpleNamespace(**base)


def test_workload_shapes_follow_baseline_configs():
    # BASELINE.json configs: cfg 1 toy split (16 000 x 64), cfg 2 640 000 x 256 per GPU (weak), cfg 3 / cfg 4 strong-scaled
    assert bench.workload_shape(_args(workload="cfg1"), 1)[:4] == (16000, 64, 4, "weak")
    for world in (1, 2, 8):
        M, d, K, scaling, scale, label = bench.workload_shape(_args(workload="cfg2"), world)
        assert (M, d, K, scaling, scale, label) == (640000, 256, 4, "weak", float(world), "cfg2")
    assert bench.workload_shape(_args(workload="cfg3"), 8)[:4] == (800000, 256, 4, "strong")
    assert bench.workload_shape(_args(workload="cfg4"), 1)[:4] == (12_800_000, 512, 8, "strong")
    assert bench.workload_shape(_args(workload="cfg4"), 8)[:4] == (6_400_000, 512, 8, "strong")      # 51.2 M rows on 8 GPUs
    M, d, K, scaling, scale, label = bench.workload_shape(_args(rows=1000, d=128, K=2), 2)
    assert (M, d, K, scaling) == (1000, 128, 2, "weak") and label.startswith("custom")


def test_config_block_names_the_workload_only():
    cfg = bench.workload_config("cfg2", 640000, 8, 256, 256, 4)
    assert cfg["workload"] == "cfg2" and cfg["rows_total"] == 5_120_000 and cfg["d_k"] == 64
    assert "larger than L2" in cfg["l2_policy"]
    assert "fit in L2" in bench.workload_config("cfg1", 16000, 1, 64, 64, 4)["l2_policy"]
    assert set(cfg) == {"workload", "rows_per_gpu", "rows_total", "d", "m", "K", "d_k", "l2_policy"}      # no run settings


@pytest.mark.parametrize("start", [0, 1, 5, 31, 32, 33, 100])
@pytest.mark.parametrize("steps", [1, 10, 16, 20, 33, 64, 200])
@pytest.mark.parametrize("every", [8, 16, 32])
def test_timed_window_carries_its_share_of_correction_steps(start, steps, every):
    align, inside = bench.correction_window(start, steps, every)
    assert 0 <= align < every
    first = start + align
    assert inside == sum(1 for i in range(first, first + steps) if i % every == 0) == int(round(steps / every))
    # the smallest such shift
    for a in range(align):
        assert sum(1 for i in range(start + a, start + a + steps) if i % every == 0) != inside


def test_modes_without_correction_steps_need_no_alignment():
    assert bench.correction_window(5, 20, 0) == (0, 0)


def test_reference_arm_prints_the_contract_line_on_cfg1():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["config"] == bench.workload_config("cfg1", 16000, 1, 64, 64, 4)          # the same dict as the product arm
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
