"""The JSON lines bench.py prints: the reference arm run here on the CPU (one bounded step of the oracle port), and
the keys of the last line measured on a B200 that is committed under profiles/ (the driver's contract)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config")


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in BASE_KEYS:
        assert k in line, k
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["impl"] == "reference" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["unit"] == baseline.get("unit", line["unit"]) and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_committed_b200_line_has_every_contract_key():
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_bench.json")))
    assert paths, "no measured bench line under profiles/"
    line = json.loads(open(paths[-1]).read().strip().splitlines()[-1])
    for k in BASE_KEYS + ("e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["value"] > 0 and line["gpu_launches"] > 0
    roof = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in roof, k
    assert roof["bound"] in ("hbm", "tensor", "l2") and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-6
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] <= line["value"] * 1.05
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
