"""bench.py contract checks that need no GPU: the reference arm's JSON line and the roofline arithmetic."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--nx", "512", "--ny", "1024"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("512x1024")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    ab = bench.algorithmic_bytes(16384, 16384)
    assert abs(ab["total"] - 8.591e9) < 2e6            # C2 8.591 GB
    assert abs(ab["rows"] + ab["cols"] - ab["total"]) < 1
    assert abs(bench.flops(16384, 16384) - 18.79e9) < 1e7   # 18.79 GF
    assert abs(bench.flops(32768, 32768) - 80.53e9) < 1e7
