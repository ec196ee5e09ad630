"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys, and the B200 arm
refuses to run without a CUDA device (there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=900)


def test_reference_arm_line():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rays/sec" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "sample" in cb
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    if build_ref.available():      # the unmodified reference (oracle/_ref) at the full 4096-ray batch: same config as the B200 arm
        assert cb["kind"] == "reference" and cb["same_config"] is True and "4096 rays/step" in cb["sample"] and "UNMODIFIED" in cb["sample"]
    else:
        assert cb["kind"] == "port" and cb["same_config"] is False
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    res = _run("--steps", "1")
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)
