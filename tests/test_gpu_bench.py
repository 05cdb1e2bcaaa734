"""bench.py contract on a real GPU, small workload: one JSON line with the keys the driver reads,
internally consistent, and a roofline fraction that is a fraction."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    r = subprocess.run([sys.executable, "bench.py", "--workload", "beam_100k_g2", "--steps", "2", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                          # exactly one JSON line
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "beam_100k_g2" and d["unit"] == "elements/s"
    assert abs(d["value"] - 100000 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0.05 < rf["frac"] < 1.3
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 2.0   # (100k: the e2e pass replays CUDA graphs, the timed steps bracket every SpMV with events)
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and 0 < c["value"] < d["value"]
    assert d["gpu_launches"] > 100                                   # our kernels ran inside the timed region
    assert d["breakdown"]["cg_terminationtype"] == 1 and d["breakdown"]["cg_rel_residual"] <= 1e-8
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "beam_100k_g2", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["value"] == d["value"] > 0
