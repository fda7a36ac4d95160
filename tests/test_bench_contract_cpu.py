"""The bench.py JSON contract: the reference arm runs for real on CPU (small video, one step), and the committed
line of our own arm (profiles/, measured on B200) is checked for the keys the driver and the judge read."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--video", "256",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "DiT denoise steps/s"
    assert d["unit"] == "steps/s" and d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/dit.py" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_line_of_our_arm_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r1_bench_1gpu_1.3B_480p.json")) as fh:
        d = json.loads(fh.read().strip().splitlines()[-1])
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["metric"] == "DiT denoise steps/s" and d["n_gpus"] == 1 and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["warmup"] >= 3
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and 0 < r["frac"] < 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    k = d["clocks"]
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert k["sm_mhz"] > 0 and not (bad & set(k["reasons"]))
