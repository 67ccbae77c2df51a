"""bench.py's reference arm and JSON contract (CPU only; the CUDA arm needs a GPU and is exercised on the box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "Msamples/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["e2e"] == {"value": j["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0,
                                           "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["value"] == j["value"] and cb["cores"] >= 1 and cb["sample"]
    import oracle
    assert cb["kind"] == ("reference" if oracle.ref_available() else "port")
    assert j["config"]["workload"].startswith("BASELINE config 2") and j["config"]["resolution"] == "1920x1080"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and not out.stdout.strip()


def test_cuda_arm_without_a_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "needs a GPU" in (out.stderr + out.stdout)
