"""bench.py prints ONE JSON line with the keys the driver reads; the reference arm works without a GPU
(CPU oracle port) and marks itself `"impl": "reference"`."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e", "cpu_baseline")


def run_bench(*args):
  res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert res.returncode == 0, res.stderr[-2000:]
  lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
  assert len(lines) == 1, lines                       # nothing but the result line on stdout
  return json.loads(lines[0])


def test_reference_arm_line_without_gpu():
  line = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
  for key in COMMON:
    assert key in line, key
  assert line["impl"] == "reference" and line["metric"] == "featurised frames/sec" and line["unit"] == "frames/s"
  assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
  assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
  assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
  assert "configs[1]" in line["config"]["workload"]


@pytest.mark.gpu
def test_native_arm_line():
  line = run_bench("--steps", "4", "--warmup", "3")
  for key in COMMON + ("gpu_launches", "clocks", "roofline"):
    assert key in line, key
  assert "impl" not in line or line["impl"] != "reference"
  assert line["n_gpus"] == 1 and line["steps"] == 4 and line["warmup"] == 3 and line["dtype"] == "f32" and line["data"] == "synthetic"
  assert line["value"] > 1000 and line["e2e"]["value"] > 1000
  assert line["e2e"]["h2d_bytes_per_step"] > 2e7 and line["e2e"]["d2h_bytes_per_step"] == 40 * 32 ** 3 * 4
  assert line["gpu_launches"] >= 4 * 4                                 # at least crop, normalisation and two tile launches per step
  roof = line["roofline"]
  assert roof["bound"] in ("hbm", "tensor") and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
  assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
  assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
  assert "larger than L2" in line["config"]["l2"]
