"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and
our arm refuses to run without a CUDA device (there is no CPU fallback to time by accident)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--points", "256", "--k", "8", "--cpu-batch", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clouds/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("point clouds/sec DGCNNSeg") and d["value"] > 0 and d["vs_baseline"] is None
    # "reference" = the unmodified reference module (staged under baseline/_ref or /root/reference), "port" = oracle
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["cpu_step_batch"] == 1 and d["config"]["same_batch_as_gpu_arm"] is False
    assert d["e2e"] == {"value": d["value"], "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_our_arm_fails_loudly_without_a_gpu():
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
