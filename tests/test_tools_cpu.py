"""Host-side tooling that needs no GPU: the CG timeline summary on a synthetic trace, and the bench lines committed
under profiles/ against the keys bench.py's contract names."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cg_timeline_on_synthetic_trace(tmp_path):
    # two ranks, 12 iterations: spmv begin (1) -> local sum (2) -> end (3) -> update begin (4) -> local (5) -> end (6)
    # -> direction begin (7) -> flags raised (9) -> wait end (11), 100 us per iteration
    steps = [(1, 0), (2, 70), (3, 72), (4, 75), (5, 85), (6, 87), (7, 90), (9, 96), (11, 98)]
    for rank in (0, 1):
        with open(tmp_path / f"stan_cg_trace_rank{rank}.csv", "w") as f:
            for k in range(12):
                for code, us in steps:
                    f.write(f"{1_000_000 + (100 * k + us) * 1000 + rank},{k + 50},{code}\n")
    r = subprocess.run([sys.executable, "tools/cg_timeline.py", str(tmp_path)], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "2 rank(s)" in out
    assert "| spmv begin | spmv local sum done | 24 | 70.0 | 70.0 |" in out
    assert "| halo wait end | spmv begin | 22 | 2.0 | 2.0 |" in out              # the gap between two iterations
    assert "mean over ranks: 100.0 us" in out


def test_committed_bench_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_10m*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r02_weak_*gpu.json")))
    assert len(files) >= 8
    for path in files:
        line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        if d.get("impl") == "reference":
            assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["value"] == d["value"]
            continue
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
            assert k in d, (path, k)
        assert d["dtype"] == "f64" and d["unit"] == "elements/s" and d["higher_is_better"] is True
        n_elem = d["config"]["n_elem"]
        assert abs(d["value"] - n_elem / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"], path
        rf = d["roofline"]
        assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0.5 < rf["frac"] < 1.1
        assert 0 < d["e2e"]["value"] <= d["value"] * 1.001 and d["e2e"]["h2d_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] > 1:
            assert d["parity"]["du"] < 1e-10 and d["parity"]["converged"]
