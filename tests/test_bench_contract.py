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


def test_bench_line_helpers_on_synthetic_numbers():
    """The pure bookkeeping of the B200 arm (kernel table, roofline entry, committed ncu traffic) runs on the CPU with made-up times:
    catches key / unit mistakes that would otherwise only show on the GPU box."""
    sys.path.insert(0, ROOT)
    import bench
    tf_peak, hbm_peak, src = bench.peaks()
    assert tf_peak > 100 and hbm_peak > 1000
    kern_ms = {"cnerf_mlp_fwd_train": 2.05, "cnerf_mlp_bwd_data": 1.55, "cnerf_mlp_bwd_weights": 2.61}
    calls = {k: 2.0 for k in kern_ms}
    for terms, tag in (({"fwd": 1, "fwd_train": 1, "chain": 1, "dw": 1}, "fwd_fp16+grad_fp16"), ({"fwd": 3, "fwd_train": 3, "chain": 3, "dw": 3}, "fwd_split+grad_split")):
        table = bench.kernel_table(kern_ms, calls, bench.N_RAYS, terms, tf_peak, hbm_peak)
        assert set(table) == set(kern_ms)
        dw = table["cnerf_mlp_bwd_weights"]
        assert dw["bytes_per_point"] == bench.DW_COLUMNS_PER_POINT * (2 if terms["dw"] == 1 else 4) and 0 < dw["hbm_frac"] < 2
        top = max(kern_ms, key=kern_ms.get)
        roof = bench.roofline_of(top, table, kern_ms, tf_peak, hbm_peak, src, 550.0, bench.N_RAYS, tag)
        assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
        assert roof["traffic"] is None or roof["traffic"] > 1e9
        roof2 = bench.roofline_of("cnerf_mlp_fwd_train", table, kern_ms, tf_peak, hbm_peak, src, 550.0, bench.N_RAYS, tag)
        assert roof2["bound"] == "tensor" and roof2["mma_per_mac"] == terms["fwd_train"] and 0 < roof2["frac"] < 1
    assert bench.ncu_traffic("cnerf_mlp_bwd_weights", "fwd_fp16+grad_fp16") > 1e10       # committed capture: 11.4 GB per step
    assert bench.ncu_traffic("cnerf_mlp_bwd_weights", "no such mode") is None
    o, d, tgt, prior, mask = bench.make_batch(64, 3)
    assert o.shape == (64, 3) and mask.shape == (64, 1) and float(d.norm(dim=-1).sub(1).abs().max()) < 1e-6


_EXIT_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import bench
dist.init_process_group("gloo")
t = torch.ones(3) * (dist.get_rank() + 1)
dist.all_reduce(t)
if dist.get_rank() == 0:
    print("SUM", t.tolist())
bench.finish_multi_gpu()          # no teardown: flush, os._exit(0)
print("NOT REACHED")
"""


def test_multi_rank_exit_tears_nothing_down(tmp_path):
    """bench.finish_multi_gpu(): both ranks of a torchrun launch leave with exit code 0 right after their last collective, the
    line printed before is not lost, and nothing after the call runs (CPU / gloo stand-in for the NCCL case)."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "exit_check.py"
    script.write_text(_EXIT_SCRIPT % ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "SUM [3.0, 3.0, 3.0]" in res.stdout and "NOT REACHED" not in res.stdout
