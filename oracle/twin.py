#!/usr/bin/env python
"""Twin harness: the UNMODIFIED reference scripts (oracle/_ref, see oracle/build_ref.py) trained twice on the same synthetic
scene -- once as they are (PyTorch eager on the GPU, the reference's own path: NP/run_nerf.py:880, NP/run_nerf_view.py:2306)
and once with the hot path swapped in by ``consistentnerf_b200.dropin`` -- and the PSNR of held-out views of both.

TEST / BENCH INFRASTRUCTURE (lives under oracle/): used by tests/, bench.py's ``quality`` and ``gpu_eager`` legs and
scripts/psnr_twin.sh; never by the product.

  python oracle/twin.py make    --kind blender|blender_view|dtu|llff --root DIR [--res N]      write a synthetic scene in the reference's format
  python oracle/twin.py run     --arm ref|repo --kind ... --root DIR --iters N    train() of the unmodified script + held-out PSNR
  python oracle/twin.py twin    --kind ... --root DIR --iters N                    make + both arms (subprocesses) -> one JSON line
  python oracle/twin.py eager   [--rays 4096] [--steps K]                          reference functions, GPU eager, workload A timing

Scenes come from an ANALYTIC teacher (a handful of coloured Gaussian density blobs) rendered by the CPU oracle's volume
renderer (oracle/nerf_oracle.py: pixel_rays + composite), so images, depth priors and held-out views are mutually consistent.
Every ``run`` is its own process: the scripts need ``torch.set_default_tensor_type('torch.cuda.FloatTensor')``, which is
process-global state.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCRIPT_OF = {"blender": "run_nerf", "dtu": "run_nerf_view", "llff": "run_nerf_view", "blender_view": "run_nerf_view"}
# sampling recipe of BASELINE.json's configs (the scene / loss flags live in the scene's config file)
FULL = ["--use_viewdirs", "--N_samples", "64", "--N_importance", "128", "--N_rand", "4096"]          # configs 2-4
CONFIG1 = ["--N_samples", "32", "--N_importance", "0", "--N_rand", "1024"]                             # config 1: coarse only, CPU plumbing
DTU_TRAIN, DTU_VAL = [25, 21, 33], [32, 24, 23, 44]                 # dtu_train[:3] / dtu_val of NP/configs/pairs.th
FERN_TRAIN, FERN_VAL = [17, 2, 7, 6, 11, 1], [12, 13, 5, 19]         # fern_train[:6] / fern_val of NP/configs/pairs.th
LLFF_H, LLFF_W, LLFF_FACTOR, LLFF_VIEWS = 378, 504, 8, 20           # factor-8 fern: train() resizes the prior depths to exactly this (NP/run_nerf_view.py:836)


# ------------------------------------------------------------------------------------------------------------------
# analytic teacher
# ------------------------------------------------------------------------------------------------------------------
class Teacher:
    """sigma(x) = sum_k a_k exp(-|x - c_k|^2 / (2 s_k^2)),  colour(x) = softmax-free blend of the blobs' colours."""

    def __init__(self, seed: int = 0, n_blobs: int = 7, extent: float = 0.9, centre=(0.0, 0.0, 0.0)):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.c = (torch.rand(n_blobs, 3, generator=g) * 2 - 1) * extent * 0.75 + torch.tensor(centre)
        self.s = 0.16 + 0.22 * torch.rand(n_blobs, generator=g) * extent
        self.a = 25.0 + 40.0 * torch.rand(n_blobs, generator=g)
        self.col = 0.08 + 0.84 * torch.rand(n_blobs, 3, generator=g)

    def raw(self, pts):
        """pts [...,3] -> raw [...,4] in the reference's convention (rgb logits, density)."""
        import torch
        d2 = ((pts[..., None, :] - self.c) ** 2).sum(-1)
        w = self.a * torch.exp(-d2 / (2 * self.s ** 2))
        sigma = w.sum(-1)
        col = (w[..., None] * self.col).sum(-2) / (sigma[..., None] + 1e-6)
        col = col.clamp(0.02, 0.98)
        return torch.cat([torch.log(col / (1 - col)), (sigma - 0.5)[..., None]], -1)

    def render(self, H, W, K, c2w, near, far, white_bkgd, n_samples=160, chunk=32768):
        """-> rgb [H,W,3], depth [H,W], acc [H,W] float32 numpy (volume rendered with the oracle's compositing)."""
        import torch
        from oracle import nerf_oracle as O
        c2w = torch.as_tensor(np.asarray(c2w), dtype=torch.float32)[:3, :4]
        ro, rd = O.pixel_rays(H, W, K, c2w)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        t = torch.linspace(0.0, 1.0, n_samples)
        z = (near * (1 - t) + far * t).expand(ro.shape[0], n_samples)
        outs = {"rgb": [], "depth": [], "acc": []}
        for i in range(0, ro.shape[0], chunk):
            pts = O.ray_points(ro[i:i + chunk], rd[i:i + chunk], z[i:i + chunk])
            c = O.composite(self.raw(pts), z[i:i + chunk], rd[i:i + chunk], None, white_bkgd)
            for k in outs:
                outs[k].append(c[k])
        return (torch.cat(outs["rgb"]).reshape(H, W, 3).numpy(), torch.cat(outs["depth"]).reshape(H, W).numpy(),
                torch.cat(outs["acc"]).reshape(H, W).numpy())


def look_at_gl(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """c2w [4,4], OpenGL camera axes (x right, y up, z backwards): the convention of the Blender / LLFF / converted DTU poses."""
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    zb = eye - target
    zb /= np.linalg.norm(zb)
    x = np.cross(up, zb)
    x /= np.linalg.norm(x)
    y = np.cross(zb, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, zb, eye
    return m


def orbit_eyes(n, radius, elev_deg, seed, jitter=0.15, az0=0.0, az_span=360.0):
    rng = np.random.RandomState(seed)
    eyes = []
    for i in range(n):
        az = math.radians(az0 + az_span * i / n + rng.uniform(-20, 20))
        el = math.radians(elev_deg + rng.uniform(-12, 12))
        r = radius * (1 + rng.uniform(-jitter, jitter) * 0.2)
        eyes.append([r * math.cos(el) * math.cos(az), r * math.cos(el) * math.sin(az), r * math.sin(el)])
    return eyes


# ------------------------------------------------------------------------------------------------------------------
# scenes in the reference's on-disk formats
# ------------------------------------------------------------------------------------------------------------------
def make_blender(root: str, res: int = 400, n_train: int = 3, n_test: int = 4, seed: int = 0, view_script: bool = False) -> dict:
    """NeRF-synthetic 'lego' stand-in (BASELINE config 2: 3 training views, white background, near 2 / far 6) under
    <root>/data/nerf_synthetic/lego, plus configs/pairs.th and (view script) the prior depths nerf_synthesic_data_depth/lego."""
    from consistentnerf_b200 import formats
    teacher = Teacher(seed)
    cam_x = 0.6911112070083618
    focal = 0.5 * res / math.tan(0.5 * cam_x)
    K = [[focal, 0, 0.5 * res], [0, focal, 0.5 * res], [0, 0, 1]]
    base = os.path.join(root, "data", "nerf_synthetic", "lego")
    eyes = orbit_eyes(n_train + 1 + n_test, 4.0, 30.0, seed)
    rng = np.random.RandomState(seed + 1)
    order = rng.permutation(len(eyes))
    eyes = [eyes[i] for i in order]
    imgs, poses, depths = [], [], []
    for e in eyes:
        c2w = look_at_gl(e)
        rgb, depth, acc = teacher.render(res, res, K, c2w, 2.0, 6.0, white_bkgd=False)
        rgba = np.concatenate([np.where(acc[..., None] > 1e-6, rgb / np.maximum(acc[..., None], 1e-6), 0.0), acc[..., None]], -1)
        imgs.append((255 * np.clip(rgba, 0, 1) + 0.5).astype(np.uint8))
        poses.append(c2w)
        depths.append(depth + (1 - acc) * 6.0)
    sl = {"train": slice(0, n_train), "val": slice(n_train, n_train + 1), "test": slice(n_train + 1, None)}
    if view_script:      # load_blender_view_data takes every split from transforms_train.json, indexed through pairs.th (NP/load_blender.py:150-180)
        splits = {"train": (imgs, poses), "val": (imgs[:1], poses[:1]), "test": (imgs[:1], poses[:1])}
        os.makedirs(os.path.join(root, "nerf_synthesic_data_depth", "lego"), exist_ok=True)
        for i, d in enumerate(depths):
            formats.write_pfm(os.path.join(root, "nerf_synthesic_data_depth", "lego", f"depth_{i:04d}.pfm"), d.astype(np.float32))
        formats.write_pairs(os.path.join(root, "configs", "pairs.th"),
                            {"lego_train": list(range(n_train)), "lego_val": list(range(n_train + 1, len(imgs)))})
    else:
        splits = {k: (imgs[v], poses[v]) for k, v in sl.items()}
    formats.write_blender_scene(base, splits, cam_x)
    cfg = os.path.join(root, "config_blender.txt")
    with open(cfg, "w") as f:
        f.write("expname = twin_lego\nbasedir = ./logs\ndatadir = ./data/nerf_synthetic/lego\ndataset_type = blender\n\n"
                "no_batching = True\nwhite_bkgd = True\nlrate_decay = 500\n\nprecrop_iters = 0\nprecrop_frac = 0.5\n\n"
                + ("half_res = False\n" if not view_script else "half_res = False\ntrain_view_num = %d\n" % n_train))
    return {"kind": "blender", "root": root, "config": cfg, "H": res, "W": res, "focal": focal, "near": 2.0, "far": 6.0}


def make_dtu(root: str, seed: int = 0, scan: str = "scan114", n_val_render: int = 2) -> dict:
    """DTU stand-in (BASELINE config 3): 49 cameras, 512 x 640, poses in mm (the loader divides by 200), three training views
    + four validation views rendered by the teacher, prior depth = teacher depth + N(0, 0.01 far); <root>/data/DTU/<scan>."""
    from consistentnerf_b200 import formats
    H, W = formats.DTU_H, formats.DTU_W
    teacher = Teacher(seed + 7, extent=0.55)
    focal = 720.0
    Kf = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], np.float64)
    near, far = 2.125, 4.525                                     # 425 mm / 200, 905 mm / 200: the real DTU range
    eyes = orbit_eyes(formats.DTU_VIEWS, 3.3, 35.0, seed, az0=-60.0, az_span=120.0)      # a forward-facing arc, like the DTU rig
    w2c_mm, views = [], {}
    rng = np.random.RandomState(seed + 3)
    for vid, e in enumerate(eyes):
        c2w_gl = look_at_gl(e)
        c2w_cv = np.concatenate([c2w_gl[:, :1], -c2w_gl[:, 1:2], -c2w_gl[:, 2:3], c2w_gl[:, 3:4]], -1)      # OpenCV axes, scene units
        c2w_mm = c2w_cv.copy()
        c2w_mm[:3, 3] *= formats.DTU_SCALE
        w2c_mm.append(np.linalg.inv(c2w_mm))
        if vid in DTU_TRAIN + DTU_VAL[:n_val_render]:      # the other views are loaded by the script but never looked at
            rgb, depth, acc = teacher.render(H, W, Kf.tolist(), c2w_gl, near, far, white_bkgd=False, n_samples=128)
            rgb = rgb + (1 - acc[..., None]) * 0.0                # black background, as the DTU photographs
            depth_full = depth + (1 - acc) * far                  # background pixels: the far plane
            prior = depth_full + rng.normal(0.0, 0.01 * far, size=depth.shape)
            views[vid] = {"image": np.clip(rgb, 0, 1), "depth": np.where(acc > 0.5, depth_full, 0.0), "prior": prior.astype(np.float32)}
    Kq = Kf.copy()
    Kq[:2] /= 4.0
    interval = (far - near) * formats.DTU_SCALE / (192 * 1.06)
    data_root = os.path.join(root, "data", "DTU")
    formats.write_dtu_scan(data_root, scan, root, views, np.stack(w2c_mm), Kq, near * formats.DTU_SCALE, interval)
    formats.write_pairs(os.path.join(root, "configs", "pairs.th"), {"dtu_train": DTU_TRAIN + [v for v in range(49) if v not in DTU_TRAIN + DTU_VAL][:13],
                                                                    "dtu_val": DTU_VAL})
    cfg = os.path.join(root, "config_dtu.txt")
    with open(cfg, "w") as f:
        f.write(f"expname = twin_dtu\nbasedir = ./logs\ndatadir = ./data/DTU/{scan}\ndataset_type = dtu\n\n"
                "no_batching = True\nlrate_decay = 500\nno_ndc = True\ntrain_view_num = 3\n\nhardmask = True\nwith_depth_loss = True\n")
    return {"kind": "dtu", "root": root, "config": cfg, "H": H, "W": W, "focal": focal, "near": near, "far": far, "scan": scan}


def make_llff(root: str, seed: int = 0, scene: str = "fern") -> dict:
    """LLFF stand-in (BASELINE config 4): 20 forward-facing cameras on a jittered grid looking at a blob cluster ~4 units down -z,
    378 x 504 at factor 8, six training + four validation views rendered by the teacher (the others are loaded, never looked at);
    <root>/data/nerf_llff_data/<scene> (poses_bounds.npy in LLFF axis order, images/, images_8/), prior depths
    <root>/nerf_llff_data_depth/<scene>/*.pfm in the units train() uses them in -- z-depth in the loader's RESCALED world
    (1 / (0.75 min bound), NP/load_llff.py:300-303; the recentring is rigid) -- and configs/pairs.th.  The config trains through
    the NDC path (no `no_ndc`: near 0 / far 1, NP/run_nerf_view.py:876-878) with the hard masks on and the official fern recipe's
    raw_noise_std = 1."""
    from consistentnerf_b200 import formats
    H, W, factor = LLFF_H, LLFF_W, LLFF_FACTOR
    teacher = Teacher(seed + 11, extent=0.9, centre=(0.0, 0.0, -4.0))
    focal = 407.5                                                    # fern: 3260 px at full resolution / 8
    K = [[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]]
    near_t, far_t = 2.0, 6.5
    rng = np.random.RandomState(seed + 5)
    rendered = set(FERN_TRAIN + FERN_VAL)
    imgs, rows, priors = [], [], []
    bd_lo, bd_hi = 2.5, 6.0
    sc = 1.0 / (bd_lo * 0.75)                                        # what load_llff_data will multiply translations and bounds with
    for i in range(LLFF_VIEWS):
        eye = [0.9 * ((i % 5) / 4.0 - 0.5) + rng.uniform(-0.05, 0.05), 0.6 * ((i // 5) / 3.0 - 0.5) + rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05)]
        c2w = look_at_gl(eye, target=(0.0, 0.0, -4.0), up=(0.0, 1.0, 0.0))
        if i in rendered:
            rgb, depth, acc = teacher.render(H, W, K, c2w, near_t, far_t, white_bkgd=False, n_samples=128)
            imgs.append((255 * np.clip(rgb, 0, 1) + 0.5).astype(np.uint8))
            prior = (depth + (1 - acc) * far_t) * sc
            priors.append((prior + rng.normal(0.0, 0.01, size=prior.shape)).astype(np.float32))
        else:
            imgs.append(np.zeros((H, W, 3), np.uint8))
            priors.append(np.full((H, W), far_t * sc, np.float32))
        llff = np.concatenate([-c2w[:3, 1:2], c2w[:3, 0:1], c2w[:3, 2:3], c2w[:3, 3:4], np.array([[H * factor], [W * factor], [focal * factor]])], 1)
        rows.append(np.concatenate([llff.reshape(-1), [bd_lo + (0.0 if i == 0 else rng.uniform(0, 0.3)), bd_hi + rng.uniform(0, 0.5)]]))
    formats.write_llff_scene(os.path.join(root, "data", "nerf_llff_data", scene), np.stack(imgs), np.stack(rows), factor=factor)
    os.makedirs(os.path.join(root, "nerf_llff_data_depth", scene), exist_ok=True)
    for i, d in enumerate(priors):
        formats.write_pfm(os.path.join(root, "nerf_llff_data_depth", scene, f"depth_{i:04d}.pfm"), d)
    formats.write_pairs(os.path.join(root, "configs", "pairs.th"),
                        {f"{scene}_train": FERN_TRAIN, f"{scene}_val": FERN_VAL, "dtu_train": list(range(16))})
    cfg = os.path.join(root, "config_llff.txt")
    with open(cfg, "w") as f:
        f.write(f"expname = twin_{scene}\nbasedir = ./logs\ndatadir = ./data/nerf_llff_data/{scene}\ndataset_type = llff\n\nfactor = {factor}\nllffhold = 8\n\n"
                "no_batching = True\nlrate_decay = 250\nraw_noise_std = 1e0\ntrain_view_num = 6\n\nhardmask = True\n")
    return {"kind": "llff", "root": root, "config": cfg, "H": H, "W": W, "focal": focal, "near": 0.0, "far": 1.0, "scene": scene, "bd_scale": sc}


def make_scene(kind: str, root: str, res: int = 400, seed: int = 0, n_train: int = 3) -> dict:
    os.makedirs(root, exist_ok=True)
    meta_path = os.path.join(root, f"scene_{kind}.json")
    if os.path.exists(meta_path):
        return json.load(open(meta_path))
    t0 = time.time()
    if kind == "blender":
        meta = make_blender(root, res=res, seed=seed, n_train=n_train)
        meta["n_train"] = n_train
    elif kind == "blender_view":
        meta = make_blender(root, res=res, seed=seed, view_script=True, n_train=n_train)
        meta["kind"] = "blender_view"
    elif kind == "dtu":
        meta = make_dtu(root, seed=seed)
    elif kind == "llff":
        meta = make_llff(root, seed=seed)
    else:
        raise ValueError(kind)
    meta["make_seconds"] = time.time() - t0
    json.dump(meta, open(meta_path, "w"))
    return meta


# ------------------------------------------------------------------------------------------------------------------
# one arm: train() of the unmodified script, then held-out PSNR through the script's own render()
# ------------------------------------------------------------------------------------------------------------------
def _heldout(meta):
    """(poses [V,4,4] c2w in the script's convention, images [V,H,W,3] float32, K, near, far) of the held-out views."""
    from consistentnerf_b200 import formats
    if meta["kind"] in ("blender", "blender_view"):
        imgs, poses, hwf, i_split = formats.load_blender_scene(os.path.join(meta["root"], "data", "nerf_synthetic", "lego"))
        if meta["kind"] == "blender":
            idx = i_split[2]
        else:
            idx = np.arange(4, len(i_split[0]))
        im = imgs[idx]
        im = im[..., :3] * im[..., -1:] + (1.0 - im[..., -1:])
        H, W, focal = hwf
        K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
        return poses[idx], im.astype(np.float32), K, 2.0, 6.0
    if meta["kind"] == "dtu":
        d = formats.load_dtu_scan(os.path.join(meta["root"], "data", "DTU"), meta["scan"], meta["root"], DTU_VAL)
        focal = float(d["K"][0, 0])
        H, W = d["images"].shape[1:3]
        K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])      # what train() builds from hwf (NP/run_nerf_view.py:961-967)
        return d["poses"], d["images"], K, float(meta["near"]), float(meta["far"])
    if meta["kind"] == "llff":
        # the poses the script trains in are the loader's: rescaled and recentred (NP/load_llff.py:288-356) -- take them from the reference loader
        import contextlib
        import io
        import load_llff                                             # oracle/_ref (on sys.path in run_arm)
        cwd = os.getcwd()
        os.chdir(meta["root"])                                       # (the loader probes ./data/midas_llff_depth relative to the cwd)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                images, poses, _bds, _rp, _it, _mono = load_llff.load_llff_data(os.path.join("data", "nerf_llff_data", meta["scene"]), LLFF_FACTOR,
                                                                                recenter=True, bd_factor=.75, spherify=False)
        finally:
            os.chdir(cwd)
        focal = float(poses[0, 2, 4])
        H, W = images.shape[1:3]
        K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
        c2w = np.tile(np.eye(4, dtype=np.float32), (len(FERN_VAL), 1, 1))
        c2w[:, :3, :4] = poses[FERN_VAL, :3, :4]
        return c2w, images[FERN_VAL].astype(np.float32), K, 0.0, 1.0  # NDC: near 0, far 1 (NP/run_nerf_view.py:876-878)
    raise ValueError(meta["kind"])


def run_arm(arm: str, kind: str, root: str, iters: int, seed: int = 0, eval_views: int = 2, eval_res_div: int = 1, extra_args=(),
            device: str = "cuda"):
    """In THIS process (must be a fresh one): train ``iters`` steps, evaluate; returns the result dict.  ``device='cpu'`` runs the
    reference arm on the host (BASELINE config 1: the scripts' plumbing without a GPU; the product has no CPU path)."""
    import torch
    extra_args = list(extra_args) or list(FULL)
    if device == "cpu" and arm != "ref":
        raise RuntimeError("the product has no CPU path; only the reference arm runs with device='cpu'")
    from consistentnerf_b200 import shims
    shims.install()
    if not os.path.isdir(REF):
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    sys.path.insert(0, REF)
    meta = json.load(open(os.path.join(root, f"scene_{kind}.json")))
    os.chdir(root)
    script = SCRIPT_OF[kind]
    import importlib
    if device == "cpu":                # run_nerf_view.py asks for the current CUDA device at import (:40) even when it then runs on the host
        torch.cuda.current_device = lambda: 0
        torch.Tensor.cuda = lambda self, *a, **k: self              # ... and get_ref_rays calls .cuda() on its constants (:596)
        torch.cuda.LongTensor = torch.LongTensor                    # ... and casts its pixel indices with .type(torch.cuda.LongTensor) (:622)
    m = importlib.import_module(script)
    patched, originals = [], {}
    if arm == "repo":
        from consistentnerf_b200 import dropin
        originals = {k: getattr(m, k) for k in dropin._RENDER_NAMES + dropin._MODEL_NAMES + dropin._VIEW_NAMES if hasattr(m, k)}
        patched = dropin.patch(m)
    if device == "cuda":
        torch.set_default_tensor_type("torch.cuda.FloatTensor")
        torch.cuda.manual_seed_all(seed)
    sync = torch.cuda.synchronize if device == "cuda" else (lambda: None)
    torch.manual_seed(seed)
    np.random.seed(seed)

    # bound the training loop (run_nerf.py hard-codes 200 001 iterations, :704) and time its steady state on the device
    stamps = {}

    def bounded_trange(a, b=None, *args, **kw):
        lo, hi = (0, a) if b is None else (a, b)
        hi = min(hi, lo + iters)
        for i in range(lo, hi):
            if i == lo + min(20, iters // 2):
                sync()
                stamps["t0"], stamps["i0"] = time.perf_counter(), i
            yield i
        sync()
        stamps["t1"], stamps["i1"] = time.perf_counter(), hi
    m.trange = bounded_trange
    # keep a handle on what the script's own create_nerf builds (the function itself runs unmodified): the held-out views are
    # rendered with exactly the networks train() optimised
    built = {}
    create_nerf = m.create_nerf

    def create_nerf_and_remember(args_):
        out_ = create_nerf(args_)
        built["train"], built["test"], built["args"] = out_[0], out_[1], args_
        return out_
    m.create_nerf = create_nerf_and_remember
    fine = "--N_importance" in extra_args and int(extra_args[extra_args.index("--N_importance") + 1]) > 0
    # run_nerf.py's checkpoint writer dereferences network_fine unconditionally (NP/run_nerf.py:801): only ask for one with a fine net
    argv = [script + ".py", "--config", meta["config"], "--i_weights", str(iters if fine else 10 ** 9), "--i_testset", str(10 ** 9),
            "--i_video", str(10 ** 9), "--i_print", str(max(1, iters // 4)), "--no_reload"] + list(extra_args)
    if script == "run_nerf_view":
        argv += ["--total_iters", str(iters), "--seed", str(seed)]
    sys.argv = argv
    t_train = time.perf_counter()
    m.train()
    sync()
    t_train = time.perf_counter() - t_train
    ms_iter = 1e3 * (stamps["t1"] - stamps["t0"]) / max(1, stamps["i1"] - stamps["i0"]) if "t0" in stamps else None

    # held-out views through the script's own render() and the render_kwargs_test its create_nerf built
    args, render_kwargs_test = built["args"], built["test"]
    ckpts = sorted(f for f in os.listdir(os.path.join(args.basedir, args.expname)) if f.endswith(".tar"))
    if fine:
        assert ckpts, "train() wrote no checkpoint"
    poses, images, K, near, far = _heldout(meta)
    render_kwargs_test.update(near=near, far=far)
    H, W = images.shape[1:3]
    psnrs, t_render = [], 0.0
    with torch.no_grad():
        for v in range(min(eval_views, len(poses))):
            Hh, Ww, Kk = H // eval_res_div, W // eval_res_div, K.copy()
            Kk[:2] /= eval_res_div
            sync()
            t0 = time.perf_counter()
            out = m.render(Hh, Ww, Kk, chunk=args.chunk, c2w=torch.Tensor(poses[v][:3, :4]), **render_kwargs_test)
            rgb = out[0]
            sync()
            t_render += time.perf_counter() - t0
            gt = images[v]
            if eval_res_div != 1:
                import cv2
                gt = cv2.resize(gt, (Ww, Hh), interpolation=cv2.INTER_AREA)
            mse = float(((rgb.cpu().numpy().astype(np.float64) - gt.astype(np.float64)) ** 2).mean())
            psnrs.append(-10.0 * math.log10(mse))
    # Render parity ON THE TRAINED WEIGHTS: the same networks, the same held-out poses, once through this package's renderer (above)
    # and once through the reference's own functions (module globals restored for the duration), so that the precision of the
    # forward is judged on a trained model and not only on the default-initialised workload A.
    same_weights = None
    if arm == "repo" and originals:
        ours = {k: getattr(m, k) for k in originals}
        try:
            for k, f in originals.items():
                setattr(m, k, f)
            RefNeRF, ref_get_embedder = originals["NeRF"], originals["get_embedder"]
            e_fn, in_ch = ref_get_embedder(args.multires, args.i_embed)
            ev_fn, in_v = ref_get_embedder(args.multires_views, args.i_embed) if args.use_viewdirs else (None, 0)
            out_ch = 5 if args.N_importance > 0 else 4

            def clone(net):
                if net is None:
                    return None
                r = RefNeRF(D=args.netdepth, W=args.netwidth, input_ch=in_ch, output_ch=out_ch, skips=[4], input_ch_views=in_v,
                            use_viewdirs=args.use_viewdirs).to(m.device)
                r.load_state_dict(net.state_dict())
                return r
            kw_ref = dict(render_kwargs_test)
            kw_ref["network_fn"], kw_ref["network_fine"] = clone(render_kwargs_test["network_fn"]), clone(render_kwargs_test.get("network_fine"))
            kw_ref["network_query_fn"] = lambda inputs, viewdirs, network_fn: m.run_network(inputs, viewdirs, network_fn, embed_fn=e_fn,
                                                                                          embeddirs_fn=ev_fn, netchunk=args.netchunk)
            diffs, psnr_r, psnr_o, all_abs = [], [], [], []
            with torch.no_grad():
                for v in range(min(2, len(poses))):
                    c2w = torch.Tensor(poses[v][:3, :4])
                    ref_rgb = m.render(H, W, K, chunk=args.chunk, c2w=c2w, **kw_ref)[0].cpu().numpy().astype(np.float64)
                    for k, f in ours.items():
                        setattr(m, k, f)
                    our_rgb = m.render(H, W, K, chunk=args.chunk, c2w=c2w, **render_kwargs_test)[0].cpu().numpy().astype(np.float64)
                    for k, f in originals.items():
                        setattr(m, k, f)
                    gt = images[v].astype(np.float64)
                    diffs.append(float(np.abs(ref_rgb - our_rgb).max()))
                    all_abs.append(np.abs(ref_rgb - our_rgb).max(-1).reshape(-1))            # per pixel: worst channel
                    psnr_r.append(-10.0 * math.log10(((ref_rgb - gt) ** 2).mean()))
                    psnr_o.append(-10.0 * math.log10(((our_rgb - gt) ** 2).mean()))
                    mse_ro = float(((ref_rgb - our_rgb) ** 2).mean())
            px = np.concatenate(all_abs)
            same_weights = {"views": len(diffs), "pixels": int(px.size), "max_abs_rgb_diff": max(diffs),
                            "abs_diff_quantiles": {q: float(np.quantile(px, float(q))) for q in ("0.5", "0.9", "0.99", "0.999", "0.9999")},
                            "frac_pixels_above_1e-4": float((px > 1e-4).mean()), "frac_pixels_above_1e-3": float((px > 1e-3).mean()),
                            "frac_pixels_above_1e-2": float((px > 1e-2).mean()), "psnr_reference_renderer": float(np.mean(psnr_r)),
                            "psnr_this_renderer": float(np.mean(psnr_o)), "delta_db": float(np.mean(psnr_o) - np.mean(psnr_r)),
                            "psnr_between_renderers": -10.0 * math.log10(max(mse_ro, 1e-30))}
        except Exception as e:      # report, do not fail the arm
            same_weights = {"error": f"{type(e).__name__}: {e}"}
        finally:
            for k, f in ours.items():
                setattr(m, k, f)
    nr = min(eval_views, len(poses)) * (H // eval_res_div) * (W // eval_res_div)
    res = {"arm": arm, "script": script + ".py", "kind": kind, "iters": iters, "patched": patched, "psnr_views": psnrs,
           "psnr": float(np.mean(psnrs)), "train_ms_per_iter": ms_iter, "train_rays_per_s": (1e3 * args.N_rand / ms_iter) if ms_iter else None, "device": device, "args": extra_args,
           "train_seconds": t_train, "render_rays_per_s": nr / t_render, "eval_views": len(psnrs), "eval_hw": [H // eval_res_div, W // eval_res_div],
           "checkpoints": ckpts, "same_weights_render_parity": same_weights,
           "fwd_precision": os.environ.get("CNERF_FWD_PRECISION", "default") if arm == "repo" else None, "grad_precision": os.environ.get("CNERF_GRAD_PRECISION", "default") if arm == "repo" else None}
    return res


def run_arm_subprocess(arm, kind, root, iters, seed=0, eval_views=2, eval_res_div=1, env=None, timeout=1800, extra_args=(), device="cuda"):
    out = os.path.join(root, f"result_{kind}_{arm}.json")
    if os.path.exists(out):
        os.remove(out)
    cmd = [sys.executable, os.path.abspath(__file__), "run", "--arm", arm, "--kind", kind, "--root", root, "--iters", str(iters), "--seed", str(seed),
           "--eval-views", str(eval_views), "--eval-res-div", str(eval_res_div), "--out", out, "--device", device] + [f"--extra={e}" for e in extra_args]
    log = open(os.path.join(root, f"log_{kind}_{arm}.txt"), "w")
    rc = subprocess.run(cmd, stdout=log, stderr=subprocess.STDOUT, env={**os.environ, **(env or {})}, timeout=timeout).returncode
    log.close()
    if rc != 0 or not os.path.exists(out):
        tail = open(os.path.join(root, f"log_{kind}_{arm}.txt")).read()[-1500:]
        return {"arm": arm, "error": f"exit {rc}", "log_tail": tail}
    return json.load(open(out))


def twin(kind, root, iters, seed=0, eval_views=2, eval_res_div=1, res=400, env_repo=None, timeout=1800, arms=("ref", "repo"), extra_args=(),
         n_train=3):
    """Scene + both arms; -> {"psnr_repo", "psnr_ref", "delta_db", "ref": {...}, "repo": {...}}."""
    make_scene(kind, root, res=res, seed=seed, n_train=n_train)
    out = {"kind": kind, "script": SCRIPT_OF[kind] + ".py", "iters": iters, "res": res}
    for arm in arms:
        out[arm] = run_arm_subprocess(arm, kind, root, iters, seed, eval_views, eval_res_div, env=env_repo if arm == "repo" else None, timeout=timeout,
                                      extra_args=extra_args)
    if all("psnr" in out.get(a, {}) for a in ("ref", "repo")):
        out.update(psnr_repo=out["repo"]["psnr"], psnr_ref=out["ref"]["psnr"], delta_db=out["repo"]["psnr"] - out["ref"]["psnr"])
    return out


# ------------------------------------------------------------------------------------------------------------------
# GPU-eager denominator: the reference's own functions on workload A (last bullet)
# ------------------------------------------------------------------------------------------------------------------
def reference_workload_a(device: str, n_rays: int, steps: int, warmup: int, mode: str, threads=None):
    """Times NP/run_nerf.py render() (+ loss + backward + Adam in train mode) built by the reference's own create_nerf on
    workload A; ``device`` 'cuda' = GPU eager (default tensor type CUDA, as the script's __main__), 'cpu' = host cores."""
    import torch
    from consistentnerf_b200 import shims
    shims.install()
    sys.path.insert(0, REF)
    import run_nerf as m
    import bench
    batches = [bench.make_batch(n_rays, b) for b in range(8)]       # drawn on the host BEFORE the default tensor type changes
    if device == "cuda":
        torch.set_default_tensor_type("torch.cuda.FloatTensor")
        m.device = torch.device("cuda")
    else:
        m.device = torch.device("cpu")
        torch.set_num_threads(threads or os.cpu_count() or 1)
    p = m.config_parser()
    args = p.parse_args(["--expname", "eager", "--basedir", "/tmp/cnerf_eager_logs", "--use_viewdirs", "--white_bkgd", "--N_samples", "64",
                         "--N_importance", "128", "--no_reload", "--dataset_type", "blender", "--chunk", "32768", "--netchunk", "65536"])
    os.makedirs("/tmp/cnerf_eager_logs/eager", exist_ok=True)
    torch.manual_seed(0)
    kw_train, kw_test, start, grad_vars, optimizer = m.create_nerf(args)
    for kw in (kw_train, kw_test):
        kw.update(near=bench.NEAR, far=bench.FAR)
    train = mode == "train"
    kw = kw_train if train else kw_test

    def step(i):
        o, d, tgt, prior, mask = batches[i % 8]
        rays = torch.stack([o, d], 0).to(m.device)
        tgt = tgt.to(m.device)
        if train:
            rgb, disp, acc, extras = m.render(1, n_rays, None, chunk=32768, rays=rays, retraw=True, **kw)
            optimizer.zero_grad()
            loss = m.img2mse(rgb, tgt) + m.img2mse(extras["rgb0"], tgt)
            loss.backward()
            optimizer.step()
            return loss
        with torch.no_grad():
            rgb, disp, acc, extras = m.render(1, n_rays, None, chunk=32768, rays=rays, **kw)
        return rgb

    def sync():
        if device == "cuda":
            torch.cuda.synchronize()
    for w in range(warmup):
        step(w)
    sync()
    t0 = time.perf_counter()
    for k in range(steps):
        r = step(k)
    float(r.reshape(-1)[0])
    sync()
    dt = (time.perf_counter() - t0) / steps
    return {"rays_per_s": n_rays / dt, "ms_per_step": dt * 1e3, "n_rays": n_rays, "steps": steps, "mode": mode, "device": device,
            "threads": torch.get_num_threads() if device == "cpu" else None,
            "what": "UNMODIFIED NP/run_nerf.py render() via its create_nerf (oracle/_ref)" + (" + img2mse x2 + backward + Adam" if train else ", no_grad")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["make", "run", "twin", "eager"])
    ap.add_argument("--kind", default="blender")
    ap.add_argument("--root", default="/tmp/cnerf_twin")
    ap.add_argument("--arm", default="repo")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--res", type=int, default=400)
    ap.add_argument("--eval-views", type=int, default=2)
    ap.add_argument("--eval-res-div", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--extra", action="append", default=[])
    ap.add_argument("--arms", default="ref,repo")
    ap.add_argument("--train-views", type=int, default=3)
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--mode", default="train")
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    if a.cmd == "make":
        res = make_scene(a.kind, a.root, res=a.res, seed=a.seed, n_train=a.train_views)
    elif a.cmd == "run":
        res = run_arm(a.arm, a.kind, a.root, a.iters, a.seed, a.eval_views, a.eval_res_div, extra_args=a.extra, device=a.device)
    elif a.cmd == "twin":
        res = twin(a.kind, a.root, a.iters, a.seed, a.eval_views, a.eval_res_div, res=a.res, extra_args=a.extra, arms=tuple(a.arms.split(",")), n_train=a.train_views)
    else:
        res = reference_workload_a(a.device, a.rays, a.steps, a.warmup, a.mode)
    text = json.dumps(res)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
