"""bench.py contract, the part that runs without a GPU: `--impl reference` times the reference's own CPU
path (oracle/_ref) and prints exactly ONE JSON line on stdout with the keys the driver reads; with no GPU to
produce the engine's generators it falls back to a matrix the reference compresses itself (`--ref-n`)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"            # what torchrun exports; the arm must undo it (re-exec)
    env["CUDA_VISIBLE_DEVICES"] = ""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-n", "4096",
                        "--steps", "2", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "HSS apply+ULV GFLOP/s" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["steps"] >= 1 and d["warmup"] == 1 and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["cores"] >= 1
    # the OMP_NUM_THREADS=1 of the caller did not survive: all host cores (or the best of the sweep) were used
    assert cb["cores"] == (os.cpu_count() or 1) or "sweep" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("HSS apply")
