"""bench.py keeps the driver's JSON contract (reference arm on CPU; the B200 arm on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-1500:]
    return json.loads(lines[0])


def test_reference_arm_line(built):
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--res", "256"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "ray_steps_per_sec_fp64" and d["unit"] == "ray-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 1e4 and d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank0_prints(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0", "--res", "64", "--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line(built):
    d = _run(["--steps", "2", "--warmup", "3", "--res", "256", "--snapshot-cells", "32", "--strong-res", "0",
              "--cpu-sample", "32"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["metric"] == "ray_steps_per_sec_fp64" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["value"] > 1e9 and d["gpu_launches"] == 2
    e = d["e2e"]
    assert e["value"] > 1e8 and e["h2d_bytes_per_step"] == 256 * 256 * 64 and e["d2h_bytes_per_step"] == 256 * 256 * 76
    rf = d["roofline"]
    assert rf["bound"] == "fp64" and rf["unit"] == "TFLOP/s" and 0.05 < rf["frac"] < 1.05 and rf["peak"] > 10
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] > 1e4 and cb["cores"] >= 1
    assert d["render"]["ms"] > 0 and d["render"]["in_domain_samples"] > 0
    ov = d["reference_scan_overhead"]
    assert ov["N"] == 10000 and 5 < ov["factor"] < 40 and abs(ov["factor"] * ov["mean_steps_per_ray"] - 10000) < 1e-6
    gp = d["generic_metric_plugin"]
    assert gp["ray_steps_per_s"] > 1e8 and gp["fp64_instr_per_ray_step"] == 2003 and 0.05 < gp["fp64_issue_frac"] < 1.05
    assert "parity_note" in d["config"] and d["roofline"]["traffic_source"].startswith("static")
