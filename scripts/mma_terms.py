"""Measured error / speed of the reduced-MMA variants of the fused forward (VERDICT r1 task 4).

Every fp32 operand is carried as fp16 hi + lo; the production forward issues a_hi*w_hi + a_hi*w_lo + a_lo*w_hi.  This script
renders workload A (4096 rays x (64 + 128)) with each SUBSET of the three products through
cnerf_debug_mlp_fwd_terms (the same kernel, a measurement instantiation), and reports against the fp64 oracle
  * max scale-relative error of rgb / depth / acc / weights over the rays whose last-sample density is decided
    (|sigma_last| >= eps, SURVEY.md section 7 hard part 2) and over the undecided bucket separately,
  * the render-only time per 4096-ray batch.
Also: the same for the training gradients (split / dw16 / fp16 backward modes) as cosine + max relative error of the flat
gradient against the fp64 oracle gradient, and the training-step time of each mode.

    python scripts/mma_terms.py [--rays 4096] > gpurun_out/mma_terms.txt
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import bench
import consistentnerf_b200 as cn
from consistentnerf_b200 import _lib, ops
from oracle import nerf_oracle as O
from util import ARCH

VARIANTS = [(7, "hi*hi + hi*lo + lo*hi (production)"), (3, "hi*hi + a_hi*w_lo (activations rounded to fp16)"),
            (5, "hi*hi + a_lo*w_hi (weights rounded to fp16)"), (1, "hi*hi only (plain fp16 operands)")]
EPS = 1e-3


def scale_rel(a, b, sel=None):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if sel is not None:
        if int(sel.sum()) == 0:
            return float("nan")
        a, b = a[sel], b[sel]
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda")
    n = args.rays
    coarse, fine = bench.make_nets(dev)
    embed_fn, _ = cn.get_embedder(10, 0)
    embeddirs_fn, _ = cn.get_embedder(4, 0)
    query = lambda i, v, f: cn.run_network(i, v, f, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
    kw = dict(network_query_fn=query, perturb=0.0, N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse,
              use_viewdirs=True, white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False, near=2.0, far=6.0)
    o, d = bench.workload_rays(n, 0)
    od = (o.to(dev), d.to(dev))

    # fp64 oracle (torch CPU; ~1 min for 4096 rays on 16 cores)
    t0 = time.time()
    pc = {k: v.detach().cpu().double() for k, v in coarse.state_dict().items()}
    pf = {k: v.detach().cpu().double() for k, v in fine.state_dict().items()}
    with torch.no_grad():
        ref = O.render_rays(O.pack_rays(o.double(), d.double(), 2.0, 6.0, True), pc, pf, ARCH, n_samples=64, n_importance=128,
                            white_bkgd=True, retraw=True)
    print(f"# fp64 oracle on {n} rays: {time.time() - t0:.1f} s", flush=True)
    sigma_last = ref["raw"][:, -1, 3]
    decided = sigma_last.abs() >= EPS
    print(f"# rays with |sigma_last| < {EPS}: {int((~decided).sum())} of {n}")

    real_fwd = ops.fused_mlp_forward
    rows = []
    for terms, name in VARIANTS:
        def fwd(packed, pts, viewdirs, terms=terms):
            nn_, S = pts.shape[0], pts.shape[1]
            raw = torch.empty((nn_, S, 4), device=pts.device, dtype=torch.float32)
            _lib.call("cnerf_debug_mlp_fwd_terms", packed.handle, _lib.ptr(pts.contiguous()), _lib.ptr(viewdirs.contiguous()), nn_, S,
                      _lib.ptr(raw), terms, _lib.stream())
            return raw
        ops.fused_mlp_forward = fwd
        try:
            with torch.no_grad():
                for _ in range(3):
                    rgb, disp, acc, depth, ex = cn.render(1, n, None, chunk=32768, rays=od, retraw=True, **kw)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    cn.render(1, n, None, chunk=32768, rays=od, **kw)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
        finally:
            ops.fused_mlp_forward = real_fwd
        got = {"rgb_map": rgb, "depth_map": depth, "acc_map": acc, "rgb0": ex["rgb0"], "depth0": ex["depth0"], "raw": ex["raw"]}
        row = {"terms": terms, "variant": name, "render_ms": ms}
        for k, v in got.items():
            row[k] = scale_rel(v, ref[k], decided)
            row[k + "_undecided"] = scale_rel(v, ref[k], ~decided)
        rows.append(row)
        print(json.dumps(row), flush=True)

    # production entry point for reference (no measurement instantiation): speed only
    with torch.no_grad():
        for _ in range(3):
            cn.render(1, n, None, chunk=32768, rays=od, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cn.render(1, n, None, chunk=32768, rays=od, **kw)
        e1.record()
        torch.cuda.synchronize()
    print(json.dumps({"variant": "cnerf_mlp_fwd (production instantiation)", "render_ms": e0.elapsed_time(e1) / 10}), flush=True)

    # ---- gradients of the three backward modes vs the fp64 oracle gradient (same rays, perturb = 0) ----
    m = min(n, 1024)
    o2, d2 = o[:m], d[:m]
    tgt = torch.rand(m, 3, generator=torch.Generator().manual_seed(5))
    names = [k for k in pc if k not in ("temp_rgb", "temp_depth", "depth_scale")]
    for p in (pc, pf):
        for k in names:
            p[k].requires_grad_(True)
    out = O.render_rays(O.pack_rays(o2.double(), d2.double(), 2.0, 6.0, True), pc, pf, ARCH, n_samples=64, n_importance=128, white_bkgd=True)
    loss = ((out["rgb_map"] - tgt.double()) ** 2).mean() + ((out["rgb0"] - tgt.double()) ** 2).mean()
    loss.backward()
    gref = torch.cat([p[k].grad.reshape(-1) for p in (pc, pf) for k in names])
    hot = [dict(net.named_parameters())[k] for net in (coarse, fine) for k in names]
    kw_t = dict(kw, perturb=0.0)
    for mode in ("split", "dw16", "fp16"):
        ops.set_grad_precision(mode)
        for p in hot:
            p.grad = None
        rgb, disp, acc, depth, ex = cn.render(1, m, None, chunk=32768, rays=(o2.to(dev), d2.to(dev)), retraw=True, **kw_t)
        l = cn.img2mse(rgb, tgt.to(dev)) + cn.img2mse(ex["rgb0"], tgt.to(dev))
        l.backward()
        g = torch.cat([p.grad.reshape(-1) for p in hot]).double().cpu()
        cos = float((g * gref).sum() / (g.norm() * gref.norm()))
        print(json.dumps({"grad_mode": mode, "rays": m, "cosine_vs_fp64": cos, "one_minus_cos": 1.0 - cos,
                          "max_rel_err": float((g - gref).abs().max() / gref.abs().max()),
                          "rel_l2_err": float((g - gref).norm() / gref.norm())}), flush=True)
    ops.set_grad_precision(ops.DEFAULT_GRAD_PRECISION)


if __name__ == "__main__":
    main()
