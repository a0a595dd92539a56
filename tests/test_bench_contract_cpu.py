"""bench.py's contract (task statement, section 4): the reference arm really runs here (it is the oracle port on host
cores) and prints one JSON line with the agreed keys; the committed GPU line of the round carries the same metric /
config plus roofline, cpu_baseline, e2e, clocks and gpu_launches."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _last_json_line(text):
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    assert lines, text
    return json.loads(lines[-1])


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr
    d = _last_json_line(r.stdout)
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "rolling_mean_rows_per_s" and d["unit"] == "rows/s"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["value"] - 100_000_000 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_gpu_line_matches_the_reference_arm_config():
    gpu = _last_json_line(open(os.path.join(ROOT, "profiles", "r2_bench_line_n1.json")).read())
    ref = _last_json_line(open(os.path.join(ROOT, "profiles", "r2_bench_reference_line.json")).read())
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "gpu_launches"} <= set(gpu)
    assert gpu["metric"] == ref["metric"] and gpu["unit"] == ref["unit"] and gpu["config"] == ref["config"]
    rf = gpu["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert gpu["gpu_launches"] >= gpu["steps"] and gpu["e2e"]["h2d_bytes_per_step"] == 800_000_000
    assert gpu["parity_ok"] is True and not set(gpu["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
