"""bench.py's GPU arm on a small configuration (BASELINE configs[1]: 4096^2 fp64, one block): the one JSON line carries
every key of the contract, device-timed throughput is plausible, and the e2e leg really moves the operands."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_gpu_arm_line_small_config():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--size", "4096", "--grid", "1", "--steps", "3", "--warmup", "3",
                          "--no-cpu-baseline", "--no-int8-split"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["unit"] == "TFLOP/s" and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert 5.0 < d["value"] < 40.0                                   # fp64 tensor peak is 37.1 TFLOP/s
    assert abs(d["value"] - 2 * 4096 ** 3 / (d["ms_per_step"] * 1e-3) / 1e12) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0.1 < r["frac"] <= 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * 4096 * 4096 * 8 and e["d2h_bytes_per_step"] == 4096 * 4096 * 8
    assert 0 < e["value"] < d["value"]                               # host buffers in and out cannot beat the resident number
    assert d["gpu_launches"] >= 3
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
