#!/usr/bin/env python
"""Benchmark of the per-ray hot path (BASELINE.json: rays/s on 4096-ray x (64 coarse + 128 fine) batches).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|render] [--impl b200|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  A *step* is one pass of the
whole hot path over one batch of 4096 synthetic rays per GPU (workload A of SURVEY.md section 8d):

  train  (default)  render() -> masked RGB + depth consistency losses (coarse and fine) -> backward ->
                    one all-reduce of the flat MLP gradient (N > 1) -> Adam step
  render            render() under no_grad with perturb=0 (novel-view path)

``value`` is timed with the batch already resident in HBM; ``e2e`` copies the batch from pinned host
memory every step and reads the loss (train) / rgb (render) back, through the same public API.
``--impl reference`` times the CPU oracle port of the reference (oracle/nerf_oracle.py) on a bounded
sample of the same workload on the host cores (the reference itself cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, N_SAMPLES, N_IMPORTANCE = 4096, 64, 128
FLOP_PER_POINT = 1_186_816
FLOP_PER_RAY_FWD = FLOP_PER_POINT * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE)
POINTS_PER_RAY = N_SAMPLES + N_SAMPLES + N_IMPORTANCE
# algorithmic FLOPs per point of each traced C-ABI call (2 x MACs of the GEMMs it evaluates, fp32-equivalent)
KERNEL_FLOPS = {
    "cnerf_mlp_fwd": FLOP_PER_POINT,                       # K2+K3 forward (inference)
    "cnerf_mlp_fwd_train": FLOP_PER_POINT,                 # forward + activation record
    "cnerf_mlp_bwd_data": 2 * (128 * 256 + 8 * 256 * 256), # dX chain: views (128->256) + 8 x (256->256)
    "cnerf_mlp_bwd_weights": 2 * (593408 - 640),           # dW of the ten GEMM layers (all but the two narrow heads)
}
# algorithmic HBM bytes per point of the weight-gradient stage: its operands G_l and X_l (fp32-equivalent, 4 B per element, read
# once): encoding pass G0+G5+E = 576 columns, eight 256x256 passes = 8 x 512, views pass G9+F+V = 416  
DW_BYTES_PER_POINT = 4 * (576 + 8 * 512 + 416)
KERNEL_NAMES = {
    "cnerf_mlp_fwd": "mlp_fused3_kernel<false> (K2+K3 forward, tcgen05)",
    "cnerf_mlp_fwd_train": "mlp_fused3_kernel<true> (K2+K3 forward + activation record, tcgen05)",
    "cnerf_mlp_bwd_data": "mlp_bwd_data3_kernel (K3b data-gradient chain, tcgen05)",
    "cnerf_mlp_bwd_weights": "mlp_bwd_weight_kernel x10 passes + reductions (K3b weight gradients, tcgen05)",
}
# kernel symbol (as ncu names it) behind each traced call, for the DRAM traffic captured with `ncu --set full` (profiles/ncu_traffic.json)
KERNEL_SYMBOL = {"cnerf_mlp_fwd": "mlp_fused3_kernel", "cnerf_mlp_fwd_train": "mlp_fused3_kernel",
                 "cnerf_mlp_bwd_data": "mlp_bwd_data3_kernel", "cnerf_mlp_bwd_weights": "mlp_bwd_weight_kernel"}


def ncu_traffic(call_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the call's kernel from the committed ncu capture, or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path))[KERNEL_SYMBOL[call_name]]["dram_bytes_per_launch"]
    except Exception:
        return None
NEAR, FAR, COEF = 2.0, 6.0, 0.2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", p["bf16_tflops"]), p["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def workload_rays(n, seed=0):
    """Workload A of SURVEY.md section 8(d): o = (0,0,4) + 0.1 N(0,1), d = normalize((0,0,-1) + 0.2 N(0,1)), near 2, far 6."""
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, 4.0]) + 0.1 * torch.randn(n, 3, generator=g)
    d = torch.tensor([0.0, 0.0, -1.0]) + 0.2 * torch.randn(n, 3, generator=g)
    return o, d / d.norm(dim=-1, keepdim=True)


def make_nets(dev):
    """Coarse and fine NeRF(D=8, W=256, 63 + 27 inputs, viewdirs) with the reference's default initialisation under seeds 0 / 1
    and the density-head bias shifted by 0.5 ("trained-like": a non-uniform sample_pdf).  Built from the
    product's own module -- the B200 arm never touches oracle/."""
    import consistentnerf_b200 as cn
    nets = []
    for seed in (0, 1):
        torch.manual_seed(seed)
        net = cn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        with torch.no_grad():
            net.alpha_linear.bias += 0.5
        nets.append(net.to(dev))
    return nets


def make_batch(n, seed):
    o, d = workload_rays(n, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    tgt = torch.rand(n, 3, generator=g)
    depth_prior = 2.5 + 3.0 * torch.rand(n, generator=g)
    mask = (torch.rand(n, 1, generator=g) > 0.3).float()      # hard mask: ~70 % consistent pixels
    return o, d, tgt, depth_prior, mask


# ----------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def roofline_of(top, kernels, kern_ms, tf_peak, hbm_peak, peak_src, step_tf):
    """Roofline entry of the dominant traced kernel: the weight-gradient stage is HBM bound, the others tensor-pipe bound."""
    k = kernels[top]
    common = {"kernel": KERNEL_NAMES[top], "traffic": ncu_traffic(top), "peak_source": peak_src, "kernel_ms_per_step": kern_ms[top],
              "mlp_step_tflops": step_tf, "mlp_step_tensor_frac": step_tf / tf_peak}
    if top == "cnerf_mlp_bwd_weights":
        return {**common, "bound": "hbm", "achieved": k["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": k["hbm_frac"],
                "note": f"algorithmic operand bytes ({DW_BYTES_PER_POINT} B per point: G_l and X_l read once, fp32-equivalent) over the "
                        f"step's {POINTS_PER_RAY * N_RAYS} points / time of the call (10 GEMM passes + reductions per network); the same "
                        f"call reaches {k['achieved_tflops']:.0f} algorithmic TFLOP/s = {k['frac']:.3f} of the tensor roofline"}
    return {**common, "bound": "tensor", "achieved": k["achieved_tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": k["frac"],
            "note": "algorithmic fp32-equivalent FLOPs of the GEMMs this call evaluates over the step's "
                    f"{POINTS_PER_RAY * N_RAYS} points; every MAC is issued as 3 fp16 MMAs (hi*hi + hi*lo + lo*hi), so 1/3 is the ceiling"}


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU baseline)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import consistentnerf_b200 as cn
    from consistentnerf_b200 import _lib
    from consistentnerf_b200.distributed import FlatGrads

    train = args.mode == "train"
    if args.grad_precision:
        cn.ops.set_grad_precision(args.grad_precision)
    grad_mode = cn.ops.grad_precision()
    dw_bytes_per_point = DW_BYTES_PER_POINT if cn.ops.GRAD_PRECISIONS[grad_mode][1] == 3 else DW_BYTES_PER_POINT // 2
    coarse, fine = make_nets(dev)
    embed_fn, _ = cn.get_embedder(10, 0)
    embeddirs_fn, _ = cn.get_embedder(4, 0)

    def query(inputs, viewdirs, network_fn):
        return cn.run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
    kw = dict(network_query_fn=query, perturb=1.0 if train else 0.0, N_importance=N_IMPORTANCE, network_fine=fine,
              N_samples=N_SAMPLES, network_fn=coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0.0,
              ndc=False, lindisp=False, near=NEAR, far=FAR)
    hot = [p for net in (coarse, fine) for n_, p in net.named_parameters() if n_ in net.spec.param_names()]
    flat = FlatGrads(hot) if train else None
    opt = torch.optim.Adam(hot, lr=5e-4, betas=(0.9, 0.999), fused=True) if train else None      # one multi-tensor kernel per step

    n_batches = 8                                              # distinct batches, rotated
    host = [tuple(x.pin_memory() for x in make_batch(N_RAYS, 100 * rank + b)) for b in range(n_batches)]
    resident = [tuple(x.to(dev) for x in hb) for hb in host]
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step(batch):
        o, d, tgt, prior, mask = batch
        if train:
            flat.zero_()
            rgb, disp, acc, depth, ex = cn.render(1, N_RAYS, None, chunk=32768, rays=(o, d), retraw=True, **kw)
            loss = (cn.masked_img_loss(rgb, tgt, mask, COEF) + cn.masked_img_loss(ex["rgb0"], tgt, mask, COEF)
                    + cn.masked_depth_loss(depth, prior, mask, FAR, COEF, include_unmasked=True)
                    + cn.masked_depth_loss(ex["depth0"], prior, mask, FAR, COEF, include_unmasked=True))
            loss.backward()
            if world > 1:
                flat.allreduce(average=True)
            opt.step()
            return loss
        with torch.no_grad():
            rgb, disp, acc, depth, ex = cn.render(1, N_RAYS, None, chunk=32768, rays=(o, d), **kw)
        return rgb

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # e2e: every step's result (loss / rgb) is copied to pinned host memory inside the timed region; the copies are asynchronous
    # (two pinned slots, an event per slot guards reuse) so the host keeps issuing the next step, as a training loop that logs
    # every i_print steps does; the last results are awaited before the clock stops
    res_slots, res_events = [None, None], [None, None]

    def read_back(r, k):
        i = k & 1
        if res_events[i] is not None:
            res_events[i].synchronize()
        if res_slots[i] is None or res_slots[i].shape != r.shape:
            res_slots[i] = torch.empty(r.shape, dtype=r.dtype, pin_memory=True)
        res_slots[i].copy_(r.detach(), non_blocking=True)
        res_events[i] = torch.cuda.Event()
        res_events[i].record()

    def timed(e2e: bool):
        for w in range(args.warmup):
            r = step(tuple(x.to(dev, non_blocking=True) for x in host[w % n_batches]) if e2e else resident[w % n_batches])
            if e2e:
                read_back(r, w)
        sync_all()
        _lib.launch_count = 0
        _lib.event_trace.clear()
        if not e2e:
            for name in KERNEL_FLOPS:
                _lib.event_trace[name] = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(args.steps):
            if e2e:
                r = step(tuple(x.to(dev, non_blocking=True) for x in host[k % n_batches]))
                read_back(r, k)                                 # loss (train) or rgb (render) back on the host
            else:
                l2_flush.zero_()                                # flush L2 between timed iterations
                step(resident[k % n_batches])
        t1.record()
        sync_all()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / args.steps

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    ms_step = timed(False)
    launches = _lib.launch_count
    traces = {k: _lib.event_trace.pop(k, []) for k in KERNEL_FLOPS}
    torch.cuda.synchronize()
    kern_ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in traces.items() if v}
    kern_calls = {k: len(v) / args.steps for k, v in traces.items() if v}
    ms_e2e = timed(True)
    clk = clocks.stop() if rank == 0 else None
    # the l2 flush memset is inside the device-timed loop; measure and subtract nothing: report as is
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    tf_peak, hbm_peak, peak_src = peaks()
    kernels = {}
    for k, ms in kern_ms.items():
        tf = KERNEL_FLOPS[k] * POINTS_PER_RAY * N_RAYS / (ms * 1e-3) / 1e12      # both launches of a step (coarse + fine)
        kernels[k] = {"kernel": KERNEL_NAMES[k], "ms_per_step": ms, "calls_per_step": kern_calls[k],
                      "achieved_tflops": tf, "frac": tf / tf_peak}
    top = max(kern_ms, key=kern_ms.get)
    achieved = kernels[top]["achieved_tflops"]
    if "cnerf_mlp_bwd_weights" in kernels:      # HBM-bound stage (ncu: 73 % DRAM throughput, 30 % tensor pipe): report both rooflines
        k = kernels["cnerf_mlp_bwd_weights"]
        gbs = dw_bytes_per_point * POINTS_PER_RAY * N_RAYS / (kern_ms["cnerf_mlp_bwd_weights"] * 1e-3) / 1e9
        k["achieved_gbs"], k["hbm_frac"] = gbs, gbs / hbm_peak
    mlp_flops_step = FLOP_PER_POINT * POINTS_PER_RAY * N_RAYS * (3 if train else 1)
    step_tf = mlp_flops_step / (ms_step * 1e-3) / 1e12
    rays_per_s = N_RAYS * world / (ms_step * 1e-3)
    e2e_rays = N_RAYS * world / (ms_e2e * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in host[0])
    d2h = 4 if train else N_RAYS * 3 * 4

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = run_reference(args, sample_rays=256, steps=16, warmup=1, quiet=True)      # a few seconds of CPU work on the box's host cores
    line = {
        "metric": "rays/sec", "value": rays_per_s, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32 (fp16x3 split on tcgen05, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"A: {N_RAYS} rays/GPU x ({N_SAMPLES} coarse + {N_IMPORTANCE} fine), viewdirs, 8x256 coarse+fine MLPs, "
                               + ("train step: render + masked rgb/depth losses + backward + grad all-reduce + Adam" if train
                                  else "render-only, no_grad, perturb=0"),
                   "mode": args.mode, "grad_precision": grad_mode, "rays_per_gpu": N_RAYS, "parallelism": f"ray-tile dp{world}",
                   "l2": "256 MiB memset between timed iterations (value); e2e streams fresh host batches"},
        "e2e": {"value": e2e_rays, "unit": "rays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline_of(top, kernels, kern_ms, tf_peak, hbm_peak, peak_src, step_tf),
        "kernels": kernels,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# reference arm: CPU oracle port of the reference path
# ----------------------------------------------------------------------------------------------
def run_reference(args, sample_rays=256, steps=None, warmup=None, quiet=False):
    from oracle import nerf_oracle as O          # the reference arm / cpu_baseline leg is the one place bench.py executes oracle/
    ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    train = args.mode == "train"
    pc, pf = O.make_params(0, sigma_bias=0.5, **ARCH), O.make_params(1, sigma_bias=0.5, **ARCH)
    names = [k for k in pc if k not in ("temp_rgb", "temp_depth", "depth_scale")]
    if train:
        for p in (pc, pf):
            for k in names:
                p[k].requires_grad_(True)
        opt = torch.optim.Adam([p[k] for p in (pc, pf) for k in names], lr=5e-4, betas=(0.9, 0.999))
    gen = torch.Generator().manual_seed(0)

    def step(i):
        o, d, tgt, prior, mask = make_batch(sample_rays, i)
        rays = O.pack_rays(o, d, NEAR, FAR, True)
        if train:
            opt.zero_grad()
            out = O.render_rays(rays, pc, pf, ARCH, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, white_bkgd=True,
                                t_rand=torch.rand(sample_rays, N_SAMPLES, generator=gen),
                                u=torch.rand(sample_rays, N_IMPORTANCE, generator=gen))
            loss = (O.masked_mse(out["rgb_map"], tgt, mask, COEF, sample_rays) + O.masked_mse(out["rgb0"], tgt, mask, COEF, sample_rays)
                    + O.masked_mse(out["depth_map"], prior, mask, COEF, sample_rays, divisor=FAR)
                    + O.masked_mse(out["depth0"], prior, mask, COEF, sample_rays, divisor=FAR))
            loss.backward()
            opt.step()
            return float(loss.detach())
        with torch.no_grad():
            out = O.render_rays(rays, pc, pf, ARCH, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, white_bkgd=True)
        return float(out["rgb_map"].sum())

    for w in range(warmup):
        step(w)
    t0 = time.perf_counter()
    for k in range(steps):
        step(k)
    dt = (time.perf_counter() - t0) / steps
    val = sample_rays / dt
    base = {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{sample_rays} rays/step x {steps} steps of workload A ({args.mode}), torch CPU fp32, {torch.get_num_threads()} threads"}
    if quiet:
        return base
    line = {"impl": "reference", "metric": "rays/sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"A: ({N_SAMPLES} coarse + {N_IMPORTANCE} fine), viewdirs, 8x256 coarse+fine MLPs, {args.mode}; "
                                   f"bounded sample of {sample_rays} rays/step", "mode": args.mode},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", choices=["train", "render"], default="train")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grad-precision", choices=["split", "dw16", "fp16"], default=None,
                    help="precision of the tensor-core backward (consistentnerf_b200.ops.GRAD_PRECISIONS); default: the package default")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        run_reference(args)
        return
    run_b200(args)


if __name__ == "__main__":
    main()
