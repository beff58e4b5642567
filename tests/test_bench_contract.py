"""bench.py prints ONE JSON line with the keys the driver reads, for both arms (reference arm on CPU; own arm on a GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
        "cpu_baseline", "e2e", "gpu_launches"}


def run_bench(args, env=None, timeout=600):
    e = dict(os.environ); e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference = the CPU oracle port with every host thread on a bounded sample (the reference itself cannot be built here)"""
    d = run_bench(["--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0"], env={"XNB_REF_BUDGET_S": "1"})
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "atom-timesteps/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] >= 1 and d["config"]["atoms"] == 256000 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    e = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], cwd=ROOT, capture_output=True, text=True, timeout=120, env=e)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


@pytest.mark.gpu
def test_own_arm_line():
    d = run_bench(["--workload", "C1", "--steps", "6", "--warmup", "3"], env={"XNB_CPU_BUDGET_S": "3"})
    assert BASE | {"roofline", "clocks"} <= set(d) and "impl" not in d or d.get("impl") != "reference"
    assert d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["traffic"] is None or rf["traffic"] > 0
    e2e = d["e2e"]
    assert 0 < e2e["value"] < d["value"] and e2e["h2d_bytes_per_step"] == 48 * d["config"]["atoms"] and e2e["d2h_bytes_per_step"] >= 72 * d["config"]["atoms"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 0 and cb["cores"] >= 1
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
