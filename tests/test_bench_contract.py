"""bench.py's reference arm on CPU: exits 0 and prints ONE JSON line with the keys the driver reads
(the GPU arm needs a device and is exercised by the driver itself)."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, AUR_BENCH_ROWS="16000", **extra_env)      # (the real arm scans the full 1M rows per step)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    p = _run({})
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["metric"].startswith("RAG queries/sec")
    assert d["config"]["rows"] == 16_000 and d["config"]["dim"] == 768 and d["config"]["k"] == 32 and d["config"]["nq"] == 256
    assert "no extrapolation" in d["cpu_baseline"]["sample"] and d["cpu_baseline"]["threads"] >= 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_nonzero_ranks_exit_quietly():
    p = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""
