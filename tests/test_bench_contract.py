"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the required keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4 and "workload" in d["config"]


def test_workload_table_matches_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.WORKLOADS["c2"][:2] == (4096, 4) and bench.WORKLOADS["c3"][:2] == (65536, 4)
    assert bench.WORKLOADS["c4"][:3] == (8192, 10, 20.0) and bench.WORKLOADS["c5"][0] == 262144
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "env-steps/sec" in base["metric"] and "env-steps/sec" in bench.METRIC
