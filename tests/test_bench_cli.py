"""bench.py as the driver runs it, without a GPU: the reference arm prints ONE JSON line with the
contract's keys (it times the unmodified reference, oracle/_ref, or the oracle port on the host
cores); the B200 arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--reads", "60000")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads_scanned_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the default workload is the metric's own shape (BASELINE.json: "d=2, 20-nt pattern")
    assert d["config"]["workload"].startswith("metric") and d["config"]["distance"] == 2 and len(d["config"]["pattern"]) == 20
    assert set(d["config"]) == {"workload", "pattern", "distance", "reads_per_gpu", "bytes_per_gpu", "line_len",
                                "l2_policy", "sharding"}


def test_reference_arm_other_ranks_do_nothing():
    r = run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
            env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run("--steps", "1", "--warmup", "0", "--reads", "1000")
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and r.stdout.strip() == ""
