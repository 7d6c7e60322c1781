"""bench.py's contract, as far as it can be exercised without a GPU: the reference arm prints one JSON line with
the required keys for every workload, and our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

import refharness as rh
from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


@pytest.mark.parametrize("workload", ["c3", "c2", "c4", "c5"])
def test_reference_arm_prints_the_contract_line(workload):
    if not rh.have_ref():
        pytest.skip("oracle/_ref/libassist_ref.so not built (needs /root/reference)")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", workload, "--cpu-sample", "4",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "0", "--n-per-gpu", "64"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
    assert not any(l.startswith("{") for l in out.stdout.splitlines())
