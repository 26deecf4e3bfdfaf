"""bench.py's output contract: ONE JSON line with the keys the driver reads.  The reference arm runs on the CPU (a small
size keeps it to seconds); the GPU arm is checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "e2e", "cpu_baseline"]


def run(*args):
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), *args], text=True, cwd=ROOT, stderr=subprocess.DEVNULL)
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run("--impl", "reference", "--format", "BC1", "--size", "256", "--steps", "1", "--warmup", "0")
    for k in BASE:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "Mtexels/s" and d["unit"] == "Mtexels/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["value"] > 0


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run("--format", "BC7", "--profile", "veryfast", "--size", "1024", "--steps", "3", "--warmup", "3")
    for k in BASE + ["gpu_launches", "clocks", "roofline"]:
        assert k in d, k
    assert "impl" not in d or d["impl"] != "reference"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["gpu_launches"] >= 3 and d["value"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 1024 * 1024 * 4 and d["e2e"]["d2h_bytes_per_step"] == 1024 * 1024
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
