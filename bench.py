#!/usr/bin/env python
"""Benchmark of the per-ray hot path (BASELINE.json: rays/s on 4096-ray x (64 coarse + 128 fine) batches; PSNR vs the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|render] [--impl b200|reference] [--scaling weak|strong]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  A *step* is one pass of the whole hot path over one
batch of synthetic rays (workload A of SURVEY.md section 8d; 4096 rays per GPU with weak scaling, 4096 rays in total with strong):

  train  (default)  render() -> masked RGB + depth consistency losses (coarse and fine; global mask counts for N > 1) -> backward
                    (the fine network's gradient all-reduce overlaps the coarse network's backward) -> Adam step
  render            render() under no_grad with perturb=0 (novel-view path)

``value`` is timed with the batch already resident in HBM; ``e2e`` copies the batch from pinned host memory every step and reads
the loss (train) / rgb (render) back, through the same public API.  The default invocation also reports, in the same JSON line:

  render        render-only throughput on workload A and on one 800 x 800 image (BASELINE config 5; row stripes + gather for N > 1)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref: NP/run_nerf.py render via its create_nerf) on the host cores, same config
  gpu_eager     the same reference code in PyTorch eager on this GPU (the scripts' own regime): the like-for-like denominator
  quality       PSNR of held-out views: the unmodified run_nerf.py train() run twice on one synthetic scene -- reference eager
                vs this package through the drop-in -- plus the committed multi-seed study (profiles/r2_psnr_twin.json)

``--impl reference`` times the reference's own CPU implementation of the same step (4096 rays per step, all host threads).
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
POINTS_PER_RAY = N_SAMPLES + N_SAMPLES + N_IMPORTANCE
FLOP_PER_RAY_FWD = FLOP_PER_POINT * POINTS_PER_RAY
# algorithmic FLOPs per point of each traced C-ABI call (2 x MACs of the GEMMs it evaluates, fp32-equivalent)
KERNEL_FLOPS = {
    "cnerf_mlp_fwd": FLOP_PER_POINT,                       # K2+K3 forward (inference)
    "cnerf_mlp_fwd_train": FLOP_PER_POINT,                 # forward + activation record
    "cnerf_mlp_bwd_data": 2 * (128 * 256 + 8 * 256 * 256), # dX chain: views (128->256) + 8 x (256->256)
    "cnerf_mlp_bwd_weights": 2 * (593408 - 640),           # dW of the ten GEMM layers (all but the two narrow heads)
}
# algorithmic HBM bytes per point of the weight-gradient stage: its operands G_l and X_l read once -- encoding pass G0+G5+E = 576
# columns, eight 256x256 passes = 8 x 512, views pass G9+F+V = 416 -- at 4 B per element (fp16 hi + lo) or 2 B (fp16)
DW_COLUMNS_PER_POINT = 576 + 8 * 512 + 416
KERNEL_NAMES = {
    "cnerf_mlp_fwd": {3: "mlp_fused3_kernel<0> (K2+K3 forward, three-term fp16 split, tcgen05)",
                      1: "mlp_fused5_kernel<0> (K2+K3 forward, fp16 operands, two tiles in flight, tcgen05)"},
    "cnerf_mlp_fwd_train": {3: "mlp_fused3_kernel<save> (forward + activation record, tcgen05)",
                            1: "mlp_fused5_kernel<1> (fp16 forward + fp16 activation record, tcgen05)"},
    "cnerf_mlp_bwd_data": {3: "mlp_bwd_data3_kernel<3 terms> (K3b data-gradient chain, tcgen05)",
                           1: "mlp_bwd_data3_kernel<1 term> (K3b data-gradient chain, tcgen05)"},
    "cnerf_mlp_bwd_weights": {3: "mlp_bwd_weight_kernel x10 passes + reductions (three-term, tcgen05 + TMA)",
                              1: "mlp_bwd_weight_kernel x10 passes + reductions (fp16 operands, tcgen05 + TMA)"},
}
NEAR, FAR, COEF = 2.0, 6.0, 0.2


def ncu_traffic(call_name, precision_tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the call's kernel from the committed ncu capture, or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
        return table[precision_tag][call_name]["dram_bytes_per_launch"]
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", p["bf16_tflops"]), p["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def workload_rays(n, seed=0):
    """Workload A of SURVEY.md section 8(d): o = (0,0,4) + 0.1 N(0,1), d = normalize((0,0,-1) + 0.2 N(0,1)), near 2, far 6."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    o = torch.tensor([0.0, 0.0, 4.0], device="cpu") + 0.1 * torch.randn(n, 3, generator=g, device="cpu")
    d = torch.tensor([0.0, 0.0, -1.0], device="cpu") + 0.2 * torch.randn(n, 3, generator=g, device="cpu")
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
    g = torch.Generator(device="cpu").manual_seed(seed + 1000)
    tgt = torch.rand(n, 3, generator=g, device="cpu")
    depth_prior = 2.5 + 3.0 * torch.rand(n, generator=g, device="cpu")
    mask = (torch.rand(n, 1, generator=g, device="cpu") > 0.3).float()      # hard mask: ~70 % consistent pixels
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


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
class Harness:
    """Everything one timed configuration needs: networks, kwargs, batches, the step function, optional CUDA graph."""

    def __init__(self, args, mode, rays_per_gpu, dev, rank, world, nets=None):
        import consistentnerf_b200 as cn
        from consistentnerf_b200.distributed import FlatGrads
        self.cn, self.args, self.mode, self.n, self.dev, self.rank, self.world = cn, args, mode, rays_per_gpu, dev, rank, world
        self.train = mode == "train"
        self.coarse, self.fine = nets if nets is not None else make_nets(dev)
        embed_fn, _ = cn.get_embedder(10, 0)
        embeddirs_fn, _ = cn.get_embedder(4, 0)

        def query(inputs, viewdirs, network_fn):
            return cn.run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
        self.kw = dict(network_query_fn=query, perturb=1.0 if self.train else 0.0, N_importance=N_IMPORTANCE, network_fine=self.fine,
                       N_samples=N_SAMPLES, network_fn=self.coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0.0,
                       ndc=False, lindisp=False, near=NEAR, far=FAR)
        groups = [[p for n_, p in net.named_parameters() if n_ in net.spec.param_names()] for net in (self.coarse, self.fine)]
        self.flat = FlatGrads(groups) if self.train else None
        if self.train and world > 1:
            self.flat.overlap_with_backward([self.coarse, self.fine])
        self.opt = torch.optim.Adam([p for g in groups for p in g], lr=5e-4, betas=(0.9, 0.999), fused=True,
                                    capturable=not args.no_graph) if self.train else None
        self.n_batches = 8
        self.host = [tuple(x.pin_memory() for x in make_batch(self.n, 100 * rank + b)) for b in range(self.n_batches)]
        self.resident = [tuple(x.to(dev) for x in hb) for hb in self.host]
        self.static = tuple(torch.empty_like(x) for x in self.resident[0])       # CUDA-graph inputs
        self.graph, self.graph_out, self.graph_err = None, None, None

    def step(self, batch):
        cn, world = self.cn, self.world
        o, d, tgt, prior, mask = batch
        if not self.train:
            with torch.no_grad():
                rgb, disp, acc, depth, ex = cn.render(1, self.n, None, chunk=32768, rays=(o, d), **self.kw)
            return rgb
        from consistentnerf_b200.distributed import global_mask_counts
        self.flat.zero_()
        gc, n_glob = None, self.n
        if world > 1:      # the means of the masked losses run over the GLOBAL batch (NP/run_nerf_view.py:1647-1648)
            gc, n_glob = global_mask_counts(mask, self.n), self.n * world
        rgb, disp, acc, depth, ex = cn.render(1, self.n, None, chunk=32768, rays=(o, d), retraw=True, **self.kw)
        loss = (cn.masked_img_loss(rgb, tgt, mask, COEF, n_rand=n_glob, global_counts=gc)
                + cn.masked_img_loss(ex["rgb0"], tgt, mask, COEF, n_rand=n_glob, global_counts=gc)
                + cn.masked_depth_loss(depth, prior, mask, FAR, COEF, n_rand=n_glob, include_unmasked=True, global_counts=gc)
                + cn.masked_depth_loss(ex["depth0"], prior, mask, FAR, COEF, n_rand=n_glob, include_unmasked=True, global_counts=gc))
        loss.backward()
        if world > 1:
            self.flat.finish()      # joins the segment reductions started from inside backward (fine first, under the coarse backward)
        self.opt.step()
        return loss

    def try_capture(self):
        """Capture one step in a CUDA graph (the step is ~90 launches = 4 ms of host enqueue: launch-bound once the kernels are fast
        or the per-GPU batch is small).  Falls back to host launches, and says so, if anything in the step is not capturable."""
        if self.args.no_graph:
            return False
        try:
            for x, src in zip(self.static, self.resident[0]):
                x.copy_(src)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self.step(self.static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.step(self.static)
            self.graph, self.graph_out = g, out
            g.replay()
            torch.cuda.synchronize()
            return True
        except Exception as e:
            self.graph, self.graph_err = None, f"{type(e).__name__}: {str(e)[:300]}"
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            return False

    def run(self, batch):
        if self.graph is None:
            return self.step(batch)
        for x, src in zip(self.static, batch):
            x.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.graph_out


def time_harness(h, args, dist, world):
    """-> (ms/step with the graph if captured, ms/step with host launches, launches, kernel ms, kernel calls, e2e ms/step)."""
    from consistentnerf_b200 import _lib
    dev = h.dev
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

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
            res_slots[i] = torch.empty(r.shape, dtype=r.dtype, device="cpu", pin_memory=True)
        res_slots[i].copy_(r.detach(), non_blocking=True)
        res_events[i] = torch.cuda.Event()
        res_events[i].record()

    def timed(e2e: bool, traced: bool):
        for w in range(args.warmup):
            r = h.run(tuple(x.to(dev, non_blocking=True) for x in h.host[w % h.n_batches]) if e2e else h.resident[w % h.n_batches])
            if e2e:
                read_back(r, w)
        sync_all()
        _lib.launch_count = 0
        _lib.event_trace.clear()
        if traced:
            for name in KERNEL_FLOPS:
                _lib.event_trace[name] = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(args.steps):
            if e2e:
                r = h.run(tuple(x.to(dev, non_blocking=True) for x in h.host[k % h.n_batches]))
                read_back(r, k)                                 # loss (train) or rgb (render) back on the host
            else:
                l2_flush.zero_()                                # flush L2 between timed iterations
                h.run(h.resident[k % h.n_batches])
        t1.record()
        sync_all()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / args.steps

    # per-kernel times and the launch count come from a pass with host launches (CUDA events around the C-ABI calls); the headline
    # replays the captured graph when there is one
    graph, h.graph = h.graph, None
    ms_eager = timed(False, True)
    launches = _lib.launch_count
    traces = {k: _lib.event_trace.pop(k, []) for k in KERNEL_FLOPS}
    torch.cuda.synchronize()
    kern_ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in traces.items() if v}
    kern_calls = {k: len(v) / args.steps for k, v in traces.items() if v}
    h.graph = graph
    ms_step = timed(False, False) if graph is not None else ms_eager
    ms_e2e = timed(True, False)
    return ms_step, ms_eager, launches, kern_ms, kern_calls, ms_e2e


def kernel_table(kern_ms, kern_calls, n_rays, terms, tf_peak, hbm_peak):
    kernels = {}
    for k, ms in kern_ms.items():
        tf = KERNEL_FLOPS[k] * POINTS_PER_RAY * n_rays / (ms * 1e-3) / 1e12      # both launches of a step (coarse + fine)
        t = terms[{"cnerf_mlp_bwd_weights": "dw", "cnerf_mlp_bwd_data": "chain", "cnerf_mlp_fwd_train": "fwd_train", "cnerf_mlp_fwd": "fwd"}[k]]
        kernels[k] = {"kernel": KERNEL_NAMES[k][t], "ms_per_step": ms, "calls_per_step": kern_calls[k], "achieved_tflops": tf,
                      "frac": tf / tf_peak, "mma_per_mac": t}
    if "cnerf_mlp_bwd_weights" in kernels:      # HBM-bound stage: report both rooflines
        k = kernels["cnerf_mlp_bwd_weights"]
        nbytes = DW_COLUMNS_PER_POINT * (4 if terms["dw"] == 3 else 2)
        gbs = nbytes * POINTS_PER_RAY * n_rays / (kern_ms["cnerf_mlp_bwd_weights"] * 1e-3) / 1e9
        k["achieved_gbs"], k["hbm_frac"], k["bytes_per_point"] = gbs, gbs / hbm_peak, nbytes
    return kernels


def roofline_of(top, kernels, kern_ms, tf_peak, hbm_peak, peak_src, step_tf, n_rays, precision_tag):
    """Roofline entry of the dominant traced kernel.  The weight-gradient stage streams its operand records from HBM (reported
    against the HBM roofline, tensor fraction beside it); the forward and the chain are tensor-pipe kernels."""
    k = kernels[top]
    common = {"kernel": k["kernel"], "traffic": ncu_traffic(top, precision_tag),
              "traffic_note": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the call's two launches of a step (fine + coarse network), "
                              "like `achieved`, which is per step of the same two launches (profiles/ncu_traffic.json)",
              "peak_source": peak_src, "kernel_ms_per_step": kern_ms[top],
              "mlp_step_tflops": step_tf, "mlp_step_tensor_frac": step_tf / tf_peak, "tensor_frac": k["frac"], "mma_per_mac": k["mma_per_mac"]}
    if top == "cnerf_mlp_bwd_weights":
        return {**common, "bound": "hbm", "achieved": k["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": k["hbm_frac"],
                "note": f"algorithmic operand bytes ({k['bytes_per_point']} B per point: G_l and X_l read once) over the step's "
                        f"{POINTS_PER_RAY * n_rays} points / time of the call (10 GEMM passes + reductions per network); the same call "
                        f"reaches {k['achieved_tflops']:.0f} algorithmic TFLOP/s = {k['frac']:.3f} of the tensor roofline"}
    return {**common, "bound": "tensor", "achieved": k["achieved_tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": k["frac"],
            "note": "algorithmic fp32-equivalent FLOPs of the GEMMs this call evaluates over the step's "
                    f"{POINTS_PER_RAY * n_rays} points; {k['mma_per_mac']} tcgen05 MMA(s) are issued per algorithmic MAC, "
                    f"so {1.0 / k['mma_per_mac']:.3f} is the ceiling of this fraction"}


def image_render_record(cn, h, dev, rank, world, dist):
    """BASELINE config 5: one 800 x 800 novel view (640 000 rays, chunk 32768); for N > 1 the rows are split in stripes and the
    RGB stripes gathered on every rank (NP/run_nerf_view.py:252-294; all-gather precedent RG/train.py:333)."""
    from consistentnerf_b200 import ops
    from consistentnerf_b200.distributed import gather_rows, shard_bounds
    H = W = 800
    focal = 0.5 * W / 0.3639702342662
    K = [[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]]
    c2w = torch.tensor([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 4.0]], device="cpu")
    kw = {k: v for k, v in h.kw.items() if k not in ("near", "far", "ndc", "use_viewdirs")}
    kw["perturb"] = 0.0

    def one():
        with torch.no_grad():
            rays = ops.image_rays(H, W, K, c2w, NEAR, FAR, True, False, dev)
            lo, hi = shard_bounds(H, rank, world)
            out = cn.batchify_rays(rays[lo * W:hi * W], 32768, **kw)
            return gather_rows(out["rgb_map"], H * W) if world > 1 else out["rgb_map"]
    for _ in range(2):
        rgb = one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    t0.record()
    for _ in range(reps):
        rgb = one()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert rgb.shape == (H * W, 3)
    return {"workload": "config 5: one 800x800 image, chunk 32768" + (f", {world} row stripes + all-gather of RGB" if world > 1 else ""),
            "ms_per_image": float(ms), "rays_per_s": H * W / (float(ms) * 1e-3), "scaling": "strong"}


def finish_multi_gpu(harnesses=None):
    """End of a multi-rank run.  The captured CUDA graphs hold NCCL kernels, and NCCL's communicator teardown waits until every
    graph that references the communicator has died (measured: after graphed N = 2 and N = 4 runs had printed their lines, the
    processes never exited -- `destroy_process_group()` / interpreter finalisation -- and the launcher had to be killed).  Every rank
    has passed its last collective and finished its device work when it gets here, so nothing is torn down at all: drain the
    device, flush, and leave with os._exit (no destructors, no communicator teardown, no graph destruction).  A timer thread does
    the same 20 s later should the synchronize ever block."""
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(20.0, lambda: os._exit(0)).start()
    try:
        torch.cuda.synchronize()      # (releases the GIL while waiting: the timer can always fire)
    finally:
        os._exit(0)


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
    from consistentnerf_b200 import ops

    if args.grad_precision:
        ops.set_grad_precision(args.grad_precision)
    if args.fwd_precision:
        ops.set_forward_precision(args.fwd_precision)
    fwd_mode, grad_mode = ops.forward_precision(), ops.grad_precision()
    chain_t, dw_t = ops.GRAD_PRECISIONS[grad_mode]
    fwd_t = ops.FWD_PRECISIONS[fwd_mode]
    terms = {"fwd": fwd_t, "fwd_train": fwd_t if dw_t == 1 else 3, "chain": chain_t, "dw": dw_t}
    precision_tag = f"fwd_{fwd_mode}+grad_{grad_mode}"
    train = args.mode == "train"
    rays_per_gpu = N_RAYS // world if args.scaling == "strong" else N_RAYS
    total_rays = rays_per_gpu * world

    h = Harness(args, args.mode, rays_per_gpu, dev, rank, world)
    graphed = h.try_capture()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    ms_step, ms_eager, launches, kern_ms, kern_calls, ms_e2e = time_harness(h, args, dist, world)
    clk = clocks.stop() if rank == 0 else None

    render_rec, hr = None, None
    if train and not args.no_render_record:      # render-only throughput, driver-visible in the same line
        hr = Harness(args, "render", rays_per_gpu, dev, rank, world, nets=(h.coarse, h.fine))
        hr.try_capture()
        r_ms, r_eager, r_launch, r_kern, r_calls, r_e2e = time_harness(hr, args, dist, world)
        img = image_render_record(cn, hr, dev, rank, world, dist)
        if rank == 0:
            tf_peak, hbm_peak, peak_src = peaks()
            tf = FLOP_PER_RAY_FWD * rays_per_gpu / (r_ms * 1e-3) / 1e12
            render_rec = {"workload": f"A render-only: {rays_per_gpu} rays/GPU x ({N_SAMPLES} + {N_IMPORTANCE}), no_grad, perturb=0",
                          "value": total_rays / (r_ms * 1e-3), "unit": "rays/s", "ms_per_step": r_ms, "cuda_graph": hr.graph is not None,
                          "ms_per_step_host_launch": r_eager,
                          "e2e": {"value": total_rays / (r_e2e * 1e-3), "ms_per_step": r_e2e,
                                  "h2d_bytes_per_step": 2 * rays_per_gpu * 12, "d2h_bytes_per_step": rays_per_gpu * 12},
                          "gpu_launches": r_launch,
                          "roofline": {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak,
                                       "mma_per_mac": fwd_t, "note": "1.2445 TFLOP of MLP work per 4096-ray batch over the whole render step"},
                          "kernels": kernel_table(r_kern, r_calls, rays_per_gpu, terms, tf_peak, hbm_peak), "image": img}
    if rank != 0:
        finish_multi_gpu([h, hr])

    tf_peak, hbm_peak, peak_src = peaks()
    kernels = kernel_table(kern_ms, kern_calls, rays_per_gpu, terms, tf_peak, hbm_peak)
    top = max(kern_ms, key=kern_ms.get)
    mlp_flops_step = FLOP_PER_POINT * POINTS_PER_RAY * rays_per_gpu * (3 if train else 1)
    step_tf = mlp_flops_step / (ms_step * 1e-3) / 1e12
    h2d = sum(x.numel() * x.element_size() for x in h.host[0])
    d2h = 4 if train else rays_per_gpu * 3 * 4

    extras = {"cpu_baseline": None}
    if world == 1 and train:
        if not args.no_cpu_baseline:
            extras["cpu_baseline"] = run_reference(args, steps=2, warmup=1, quiet=True)      # ~12 s of CPU work on the box's host cores
        if not args.no_gpu_eager:
            extras["gpu_eager"] = gpu_eager_record()
        if not args.no_quality:
            extras["quality"] = quality_record(args)
    line = {
        "metric": "rays/sec", "value": total_rays / (ms_step * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": ("fp32-equivalent forward (fp16x3 split on tcgen05, fp32 accumulate)" if fwd_t == 3 else "fp16 operands on tcgen05, fp32 accumulate")
                 + f"; gradients: {grad_mode}",
        "data": "synthetic",
        "config": {"workload": f"A: {rays_per_gpu} rays/GPU x ({N_SAMPLES} coarse + {N_IMPORTANCE} fine), viewdirs, 8x256 coarse+fine MLPs, "
                               + ("train step: render + masked rgb/depth losses + backward + grad all-reduce + Adam" if train
                                  else "render-only, no_grad, perturb=0"),
                   "mode": args.mode, "forward_precision": fwd_mode, "grad_precision": grad_mode, "rays_per_gpu": rays_per_gpu,
                   "global_rays": total_rays, "parallelism": f"ray-tile dp{world}", "cuda_graph": bool(graphed),
                   "cuda_graph_error": h.graph_err, "ms_per_step_host_launch": ms_eager,
                   "l2": "256 MiB memset between timed iterations (value); e2e streams fresh host batches"},
        "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline_of(top, kernels, kern_ms, tf_peak, hbm_peak, peak_src, step_tf, rays_per_gpu, precision_tag),
        "kernels": kernels,
        "render": render_rec,
        **extras,
    }
    print(json.dumps(line))
    if world > 1:
        finish_multi_gpu([h, hr])


# ----------------------------------------------------------------------------------------------
# legs that run the reference (oracle/_ref) -- the only places bench.py executes anything under oracle/
# ----------------------------------------------------------------------------------------------
def _twin_cli(argv, timeout, env=None):
    """Run oracle/twin.py in its own process (the reference needs process-global default-tensor-type state); -> dict."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "twin.py")] + argv
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env={**os.environ, **(env or {})})
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"error": f"exit {res.returncode}: {res.stderr[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def ref_available():
    from oracle import build_ref
    return build_ref.available()


def gpu_eager_record():
    """The reference's own code in PyTorch eager on this GPU, workload A: the like-for-like denominator."""
    if not ref_available():
        return {"unavailable": "oracle/_ref not populated"}
    tr = _twin_cli(["eager", "--mode", "train", "--steps", "5", "--warmup", "2", "--rays", str(N_RAYS)], 300)
    rd = _twin_cli(["eager", "--mode", "render", "--steps", "5", "--warmup", "2", "--rays", str(N_RAYS)], 300)
    return {"what": "UNMODIFIED NP/run_nerf.py render() (+ img2mse x2 + backward + Adam) under torch.set_default_tensor_type('torch.cuda.FloatTensor'), "
                    "fp32 eager, same GPU, 4096 rays x (64 + 128)",
            "train": {k: tr.get(k) for k in ("rays_per_s", "ms_per_step", "error") if k in tr},
            "render": {k: rd.get(k) for k in ("rays_per_s", "ms_per_step", "error") if k in rd}}


def quality_record(args):
    """PSNR half of the metric: unmodified run_nerf.py train() on one synthetic Blender-format scene, reference eager vs drop-in."""
    committed = None
    try:
        committed = json.load(open(os.path.join(ROOT, "profiles", "r2_psnr_twin.json")))["summary"]
    except Exception:
        pass
    if not ref_available():
        return {"unavailable": "oracle/_ref not populated", "committed_study": committed}
    from consistentnerf_b200 import ops
    root = "/tmp/cnerf_bench_twin"
    env = {"CNERF_FWD_PRECISION": ops.forward_precision(), "CNERF_GRAD_PRECISION": ops.grad_precision()}
    t0 = time.time()
    res = _twin_cli(["twin", "--kind", "blender", "--root", root, "--iters", str(args.quality_iters), "--res", "200", "--eval-views", "4", "--train-views", "8"], 900, env=env)
    out = {"what": f"UNMODIFIED run_nerf.py train() x {args.quality_iters} iters (N_rand 4096, 64 + 128 samples, 8 training views 200x200, white bkgd) -- reference "
                   "GPU eager vs this package through consistentnerf_b200.dropin -- PSNR on 4 held-out views; live, one seed (the seed-to-seed "
                   "spread of either arm is in committed_study)",
           "seconds": time.time() - t0, "committed_study": committed}
    if "psnr_repo" in res:
        out.update(psnr_repo=res["psnr_repo"], psnr_ref=res["psnr_ref"], delta_db=res["delta_db"],
                   train_ms_per_iter_repo=res["repo"]["train_ms_per_iter"], train_ms_per_iter_ref=res["ref"]["train_ms_per_iter"],
                   script_rays_per_s_repo=res["repo"]["train_rays_per_s"], script_rays_per_s_ref=res["ref"]["train_rays_per_s"])
    else:
        out["error"] = {a: (res.get(a) or {}).get("error") or (res.get(a) or {}).get("log_tail", "")[-300:] for a in ("ref", "repo")} if "ref" in res else res
    return out


def run_reference(args, steps=None, warmup=None, quiet=False):
    """Reference arm / cpu_baseline leg: the reference's own CPU implementation of the step on the host cores.  With oracle/_ref
    (the unmodified NP/run_nerf.py, shipped by oracle/build_ref.py) it is the reference itself at the full 4096 rays per step
    (kind "reference"); without it, the oracle port on a bounded 256-ray sample (kind "port")."""
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    cores = os.cpu_count() or 1
    mode = args.mode
    if ref_available():
        # each step = the full 4096-ray batch (~4 s for a training step on 16 cores); the count is capped so that the run ends within minutes
        steps_run, warm_run = min(steps, 24), min(warmup, 1)
        r = _twin_cli(["eager", "--device", "cpu", "--mode", mode, "--steps", str(steps_run), "--warmup", str(warm_run), "--rays", str(N_RAYS)], 1500)
        if "rays_per_s" in r:
            val, dt = r["rays_per_s"], r["ms_per_step"] * 1e-3
            base = {"value": val, "unit": "rays/s", "cores": cores, "kind": "reference", "same_config": True,
                    "sample": f"{N_RAYS} rays/step x {steps_run} steps of workload A ({mode}): UNMODIFIED NP/run_nerf.py render() via its create_nerf "
                              f"(oracle/_ref), torch CPU fp32, {r.get('threads')} threads"}
            if quiet:
                return base
            line = {"impl": "reference", "metric": "rays/sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps_run,
                    "steps_requested": steps, "warmup": warm_run, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
                    "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                    "config": {"workload": f"A: {N_RAYS} rays x ({N_SAMPLES} coarse + {N_IMPORTANCE} fine), viewdirs, 8x256 coarse+fine MLPs, {mode}", "mode": mode},
                    "cpu_baseline": base, "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            print(json.dumps(line))
            return base
    return run_reference_port(args, sample_rays=256, steps=min(steps, 16), warmup=min(warmup, 1), quiet=quiet)


def run_reference_port(args, sample_rays=256, steps=None, warmup=None, quiet=False):
    from oracle import nerf_oracle as O
    ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
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
    base = {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "same_config": False,
            "sample": f"{sample_rays} rays/step x {steps} steps of workload A ({args.mode}), oracle port (oracle/_ref not shipped), "
                      f"torch CPU fp32, {torch.get_num_threads()} threads"}
    if quiet:
        return base
    line = {"impl": "reference", "metric": "rays/sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"A: ({N_SAMPLES} coarse + {N_IMPORTANCE} fine), viewdirs, 8x256 coarse+fine MLPs, {args.mode}; "
                                   f"bounded sample of {sample_rays} rays/step", "mode": args.mode},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", choices=["train", "render"], default="train")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: 4096 rays per GPU; strong: 4096 rays in total (how the reference trains), 4096 / N per GPU")
    ap.add_argument("--grad-precision", choices=["split", "dw16", "fp16"], default=None,
                    help="precision of the tensor-core backward (consistentnerf_b200.ops.GRAD_PRECISIONS); default: the package default")
    ap.add_argument("--fwd-precision", choices=["split", "fp16"], default=None,
                    help="precision of the fused forward (consistentnerf_b200.ops.FWD_PRECISIONS); default: the package default")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-quality", action="store_true")
    ap.add_argument("--no-render-record", action="store_true")
    ap.add_argument("--quality-iters", type=int, default=300)
    ap.add_argument("--quick", action="store_true", help="shorthand for --no-cpu-baseline --no-gpu-eager --no-quality --no-render-record")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_gpu_eager = args.no_quality = args.no_render_record = True
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        run_reference(args)
        return
    run_b200(args)


if __name__ == "__main__":
    main()
