"""On-disk formats of the hot path's callers -- host-side, no kernels.

* checkpoints: the ``{:06d}.tar`` files train() writes with torch.save (NP/run_nerf_view.py:2004-2015) and create_nerf reloads
  (NP/run_nerf_view.py:336-363): a dict with ``global_step``, ``network_fn_state_dict``, ``network_fine_state_dict``,
  ``optimizer_state_dict``.  On reload the reference overwrites the three soft-mask scalars (``temp_rgb``, ``temp_depth``,
  ``depth_scale``) with 0.1 in both networks (:350-355) and does NOT restore the optimizer state (:348, commented out).
* PFM depth maps: the prior depths of the Blender / DTU loaders (NP/load_blender.py:93-128 ``read_pfm``): ``Pf``/``PF`` header,
  ``width height``, a scale whose sign is the endianness, rows stored bottom-up.
* ``metrics.txt``: ``PSNR: x\nSSIM: y\nLPIPS: z`` without a trailing newline (NP/run_nerf_view.py:2078-2087).
"""
from __future__ import annotations

import os
import re
from typing import Optional, Tuple

import numpy as np
import torch

__all__ = ["save_checkpoint", "load_checkpoint", "latest_checkpoint", "read_pfm", "write_pfm", "write_metrics", "read_metrics",
           "write_pairs", "write_blender_scene", "load_blender_scene", "write_dtu_cam", "read_dtu_cam", "write_dtu_scan",
           "load_dtu_scan", "write_llff_scene", "load_llff_scene"]

_SOFTMASK_SCALARS = ("temp_rgb", "temp_depth", "depth_scale")


def save_checkpoint(path: str, global_step: int, network_fn, network_fine=None, optimizer=None) -> str:
    """Write a reference-format checkpoint (NP/run_nerf_view.py:2004-2015); returns ``path``."""
    ckpt = {"global_step": int(global_step), "network_fn_state_dict": network_fn.state_dict()}
    if network_fine is not None:
        ckpt["network_fine_state_dict"] = network_fine.state_dict()
    if optimizer is not None:
        ckpt["optimizer_state_dict"] = optimizer.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(ckpt, path)
    return path


def latest_checkpoint(directory: str) -> Optional[str]:
    """The file create_nerf would pick: the last of the sorted names containing 'tar' (NP/run_nerf_view.py:340-346)."""
    names = [f for f in sorted(os.listdir(directory)) if "tar" in f]
    return os.path.join(directory, names[-1]) if names else None


def load_checkpoint(path: str, network_fn, network_fine=None, optimizer=None, reference_reload_semantics: bool = True,
                    map_location="cpu") -> int:
    """Load a reference-format checkpoint into the given modules; returns ``global_step`` (the training loop's ``start``).

    ``reference_reload_semantics`` reproduces create_nerf (NP/run_nerf_view.py:350-363): the soft-mask scalars of both networks
    are reset to 0.1 and the optimizer state is left untouched.  With False the file is restored verbatim (and the optimizer
    too, when given)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    for key, net in (("network_fn_state_dict", network_fn), ("network_fine_state_dict", network_fine)):
        if net is None:
            continue
        sd = dict(ckpt[key])
        if reference_reload_semantics:
            for name in _SOFTMASK_SCALARS:
                sd[name] = torch.tensor([0.1])
        net.load_state_dict(sd)
    if optimizer is not None and not reference_reload_semantics and "optimizer_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return int(ckpt["global_step"])


def read_pfm(filename: str) -> Tuple[np.ndarray, float]:
    """(data [H,W] or [H,W,3] float32 with row 0 at the TOP, scale) -- NP/load_blender.py:93-128."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header not in ("PF", "Pf"):
            raise ValueError("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise ValueError("Malformed PFM header.")
        width, height = int(m.group(1)), int(m.group(2))
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.frombuffer(f.read(), dtype=endian + "f4")
    shape = (height, width, 3) if header == "PF" else (height, width)
    return np.flipud(np.reshape(data, shape)).astype(np.float32), abs(scale)


def write_pfm(filename: str, image: np.ndarray, scale: float = 1.0) -> None:
    """Inverse of read_pfm: little-endian float32, rows bottom-up, negative scale in the header."""
    image = np.asarray(image, dtype=np.float32)
    if image.ndim == 3 and image.shape[2] == 3:
        header = "PF"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        header, image = "Pf", image.reshape(image.shape[0], image.shape[1])
    else:
        raise ValueError("PFM holds [H,W], [H,W,1] or [H,W,3] images")
    with open(filename, "wb") as f:
        f.write(f"{header}\n{image.shape[1]} {image.shape[0]}\n{-abs(float(scale))}\n".encode("utf-8"))
        np.flipud(image).astype("<f4").tofile(f)


def write_metrics(path: str, psnr, ssim, lpips) -> None:
    """metrics.txt exactly as the reference writes it (NP/run_nerf_view.py:2078-2087)."""
    with open(path, "w") as f:
        f.write(f"PSNR: {psnr}\n")
        f.write(f"SSIM: {ssim}\n")
        f.write(f"LPIPS: {lpips}")


def read_metrics(path: str) -> dict:
    out = {}
    for line in open(path).read().splitlines():
        k, _, v = line.partition(":")
        m = re.search(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|nan|inf", v)      # first numeric token: 'tensor(0.12, device='cuda:0', grad_fn=...)'
        out[k.strip()] = float(m.group(0)) if m else float("nan")
    return out


# ------------------------------------------------------------------------------------------------------------------
# Dataset directories the reference's loaders read: writers + native loaders.
# The writers produce exactly the layout NP/load_blender.py, NP/load_dtu.py and NP/load_llff.py expect, so that the UNMODIFIED
# scripts train on scenes written here (oracle/twin.py); the loaders return the same arrays without imageio / the reference.
# ------------------------------------------------------------------------------------------------------------------
def _write_png(path: str, img: np.ndarray) -> None:
    import cv2
    img = np.asarray(img)
    if img.dtype != np.uint8:
        img = (255.0 * np.clip(img, 0.0, 1.0) + 0.5).astype(np.uint8)
    if img.ndim == 3 and img.shape[2] == 3:
        img = cv2.cvtColor(img, cv2.COLOR_RGB2BGR)
    elif img.ndim == 3 and img.shape[2] == 4:
        img = cv2.cvtColor(img, cv2.COLOR_RGBA2BGRA)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if not cv2.imwrite(path, img):
        raise IOError(f"cannot write {path}")


def _read_png(path: str) -> np.ndarray:
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3 and img.shape[2] == 3:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    elif img.ndim == 3 and img.shape[2] == 4:
        img = cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA)
    return img


def write_pairs(path: str, table: dict) -> None:
    """``configs/pairs.th``: {"<scene>_train": [...], "<scene>_val": [...], "dtu_train": [...], ...} of plain int lists
    (NP/run_nerf_view.py:867-870,942-945; NP/load_blender.py:174-176).  Plain lists load under torch.load's weights_only default."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({k: [int(i) for i in v] for k, v in table.items()}, path)


def write_blender_scene(basedir: str, splits: dict, camera_angle_x: float) -> None:
    """NeRF-synthetic layout (NP/load_blender.py:38-60): ``transforms_{train,val,test}.json`` with ``camera_angle_x`` and
    ``frames[*].{file_path, transform_matrix}``, images at ``<file_path>.png`` (RGBA, 8 bit).
    ``splits[name] = (images [N,H,W,4] uint8 or float in [0,1], c2w [N,4,4])``."""
    import json
    for name in ("train", "val", "test"):
        imgs, poses = splits[name]
        frames = []
        for i, (im, c2w) in enumerate(zip(imgs, poses)):
            rel = f"./{name}/r_{i}"
            _write_png(os.path.join(basedir, name, f"r_{i}.png"), im)
            frames.append({"file_path": rel, "rotation": 0.0, "transform_matrix": np.asarray(c2w, dtype=np.float64).tolist()})
        with open(os.path.join(basedir, f"transforms_{name}.json"), "w") as f:
            json.dump({"camera_angle_x": float(camera_angle_x), "frames": frames}, f, indent=1)


def load_blender_scene(basedir: str, testskip: int = 1):
    """-> (imgs [N,H,W,4] float32 in [0,1], poses [N,4,4] float32, [H, W, focal], [i_train, i_val, i_test]) -- the values of
    NP/load_blender.py:38-76 (before the half_res resize)."""
    import json
    all_imgs, all_poses, counts = [], [], [0]
    for s in ("train", "val", "test"):
        with open(os.path.join(basedir, f"transforms_{s}.json")) as f:
            meta = json.load(f)
        skip = 1 if (s == "train" or testskip == 0) else testskip
        imgs = [_read_png(os.path.join(basedir, fr["file_path"] + ".png")) for fr in meta["frames"][::skip]]
        poses = [np.array(fr["transform_matrix"]) for fr in meta["frames"][::skip]]
        all_imgs.append((np.array(imgs) / 255.0).astype(np.float32))
        all_poses.append(np.array(poses).astype(np.float32))
        counts.append(counts[-1] + len(imgs))
    imgs, poses = np.concatenate(all_imgs, 0), np.concatenate(all_poses, 0)
    H, W = imgs[0].shape[:2]
    focal = 0.5 * W / np.tan(0.5 * float(meta["camera_angle_x"]))
    return imgs, poses, [H, W, focal], [np.arange(counts[i], counts[i + 1]) for i in range(3)]


def write_dtu_cam(path: str, w2c: np.ndarray, intrinsic: np.ndarray, depth_min: float, depth_interval: float) -> None:
    """MVSNet camera file as parsed by read_cam_file (NP/load_dtu.py:131-143): 'extrinsic' + 4 rows, blank, 'intrinsic' + 3 rows,
    blank, '<depth_min> <depth_interval>'."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write("extrinsic\n")
        for r in np.asarray(w2c, dtype=np.float64).reshape(4, 4):
            f.write(" ".join(repr(float(v)) for v in r) + " \n")
        f.write("\nintrinsic\n")
        for r in np.asarray(intrinsic, dtype=np.float64).reshape(3, 3):
            f.write(" ".join(repr(float(v)) for v in r) + " \n")
        f.write(f"\n{float(depth_min)!r} {float(depth_interval)!r}\n")


def read_dtu_cam(path: str):
    """-> (intrinsic [3,3], w2c [4,4], [depth_min, depth_max]) with depth_max = depth_min + 192 * 1.06 * interval (NP/load_dtu.py:131-143)."""
    with open(path) as f:
        lines = [ln.rstrip() for ln in f.readlines()]
    ext = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intr = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    dmin, dint = (float(v) for v in lines[11].split()[:2])
    return intr, ext, [dmin, dmin + dint * 192 * 1.06]


DTU_VIEWS, DTU_SCALE, DTU_H, DTU_W = 49, 200.0, 512, 640


def write_dtu_scan(root: str, scan: str, prior_root: str, views: dict, w2c_mm: np.ndarray, intrinsic_quarter: np.ndarray,
                   depth_min_mm: float, depth_interval_mm: float, light_idx: int = 3) -> None:
    """The files load_dtu_data reads for ``--datadir <root>/<scan>`` (NP/load_dtu.py:188-217), 49 views:
      <root>/Rectified/<scan>_train/rect_%03d_<light>_r5000.png   512 x 640 RGB           (view id + 1)
      <root>/Depths/Cameras/train/%08d_cam.txt                    w2c in mm, intrinsics at 1/4 resolution (loader multiplies by 4)
      <root>/Depths/<scan>/depth_map_%04d.pfm                     1200 x 1600 mm; loader: half size -> crop [44:556, 80:720] -> / 200
      <prior_root>/nerf_dtu_data_depth/<scan>/depth_%04d.pfm      512 x 640 prior depth in scene units (mm / 200); prior_root = the CWD
    ``views[vid] = dict(image [512,640,3], depth [512,640] scene units, prior [512,640] scene units)``; views not listed get a
    black image and zero depths (hard-linked to one file each: the loader reads all 49 although a run touches only 7)."""
    blank_img = os.path.join(root, f"Rectified/{scan}_train/_blank.png")
    blank_gt = os.path.join(root, f"Depths/{scan}/_blank.pfm")
    blank_prior = os.path.join(prior_root, f"nerf_dtu_data_depth/{scan}/_blank.pfm")
    _write_png(blank_img, np.zeros((DTU_H, DTU_W, 3), np.uint8))
    os.makedirs(os.path.dirname(blank_gt), exist_ok=True)
    os.makedirs(os.path.dirname(blank_prior), exist_ok=True)
    write_pfm(blank_gt, np.zeros((1200, 1600), np.float32))
    write_pfm(blank_prior, np.zeros((DTU_H, DTU_W), np.float32))

    def link(src, dst):
        if os.path.exists(dst):
            os.remove(dst)
        os.link(src, dst)

    for vid in range(DTU_VIEWS):
        img_p = os.path.join(root, f"Rectified/{scan}_train/rect_{vid + 1:03d}_{light_idx}_r5000.png")
        gt_p = os.path.join(root, f"Depths/{scan}/depth_map_{vid:04d}.pfm")
        prior_p = os.path.join(prior_root, f"nerf_dtu_data_depth/{scan}/depth_{vid:04d}.pfm")
        write_dtu_cam(os.path.join(root, f"Depths/Cameras/train/{vid:08d}_cam.txt"), w2c_mm[vid], intrinsic_quarter, depth_min_mm,
                      depth_interval_mm)
        v = views.get(vid)
        if v is None:
            link(blank_img, img_p); link(blank_gt, gt_p); link(blank_prior, prior_p)
            continue
        _write_png(img_p, v["image"])
        big = np.zeros((1200, 1600), np.float32)              # the loader halves (nearest) and crops: place the map so that it survives
        big[88:1112:2, 160:1440:2] = np.asarray(v["depth"], np.float32) * DTU_SCALE
        big[89:1112:2, 160:1440:2] = big[88:1112:2, 160:1440:2]
        big[88:1112, 161:1440:2] = big[88:1112, 160:1440:2]
        write_pfm(gt_p, big)
        write_pfm(prior_p, np.asarray(v["prior"], np.float32))


def load_dtu_scan(root: str, scan: str, prior_root: str, view_ids, light_idx: int = 3):
    """Native loader of the views ``view_ids`` of a DTU scan: dict(images [V,512,640,3] float32, poses [V,4,4] c2w in the
    reference's convention (OpenGL axes, translation / 200; NP/load_dtu.py:199-204), bds [V,2], K [3,3] at full resolution,
    priors [V,512,640], depths [V,512,640])."""
    import cv2
    imgs, poses, bds, priors, depths = [], [], [], [], []
    K = None
    for vid in view_ids:
        intr, w2c, nf = read_dtu_cam(os.path.join(root, f"Depths/Cameras/train/{vid:08d}_cam.txt"))
        intr[:2] *= 4
        K = intr
        c2w = np.linalg.inv(w2c)
        c2w[:3, 3] *= 1.0 / DTU_SCALE
        poses.append(np.concatenate([c2w[:, :1], -c2w[:, 1:2], -c2w[:, 2:3], c2w[:, 3:4]], -1))
        imgs.append(_read_png(os.path.join(root, f"Rectified/{scan}_train/rect_{vid + 1:03d}_{light_idx}_r5000.png")).astype(np.float32) / 255.0)
        priors.append(read_pfm(os.path.join(prior_root, f"nerf_dtu_data_depth/{scan}/depth_{vid:04d}.pfm"))[0])
        d = read_pfm(os.path.join(root, f"Depths/{scan}/depth_map_{vid:04d}.pfm"))[0]
        d = cv2.resize(d, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_NEAREST)[44:556, 80:720]
        depths.append(d / DTU_SCALE)
        bds.append([nf[0] / DTU_SCALE, nf[1] / DTU_SCALE])
    return {"images": np.stack(imgs), "poses": np.stack(poses).astype(np.float32), "bds": np.array(bds, np.float32), "K": K,
            "priors": np.stack(priors), "depths": np.stack(depths)}


def write_llff_scene(basedir: str, images: np.ndarray, poses_bounds: np.ndarray, factor: int = 8) -> None:
    """LLFF layout (NP/load_llff.py:60-110): ``poses_bounds.npy`` [N,17] = 3x5 [R|t|hwf] (LLFF axis order: down, right, back)
    flattened + [near, far]; ``images/`` (only its first file's shape and the file count are read) and ``images_<factor>/`` holding
    the images that are actually loaded.  ``images`` [N,h,w,3] are the FACTOR-reduced images; the hwf column must describe the
    full resolution (h * factor, w * factor, focal * factor): the loader overwrites h, w with the loaded shape and divides the
    focal by the factor."""
    poses_bounds = np.asarray(poses_bounds, dtype=np.float64)
    if poses_bounds.shape != (len(images), 17):
        raise ValueError("poses_bounds must be [N,17]")
    os.makedirs(basedir, exist_ok=True)
    np.save(os.path.join(basedir, "poses_bounds.npy"), poses_bounds)
    for i, im in enumerate(images):
        _write_png(os.path.join(basedir, f"images_{factor}", f"image{i:03d}.png"), im)
        _write_png(os.path.join(basedir, "images", f"image{i:03d}.png"), np.zeros((2, 2, 3), np.uint8))      # placeholder: never decoded at full size


def load_llff_scene(basedir: str, factor: int = 8):
    """-> (poses [3,5,N], bds [2,N], imgs [h,w,3,N] float in [0,1]) as NP/load_llff.py:_load_data returns them (:60-115)."""
    arr = np.load(os.path.join(basedir, "poses_bounds.npy"))
    poses = arr[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0]).copy()
    bds = arr[:, -2:].transpose([1, 0])
    d = os.path.join(basedir, f"images_{factor}")
    files = [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.lower().endswith(("jpg", "png"))]
    if poses.shape[-1] != len(files):
        raise ValueError(f"Mismatch between imgs {len(files)} and poses {poses.shape[-1]}")
    imgs = np.stack([_read_png(f)[..., :3] / 255.0 for f in files], -1)
    poses[:2, 4, :] = np.array(imgs.shape[:2]).reshape([2, 1])
    poses[2, 4, :] = poses[2, 4, :] * 1.0 / factor
    return poses, bds, imgs


def softmask_path(root: str, dataset_type: str, scene: str, index: int, top_k: int = 30) -> str:
    """Where train(--softmask) looks for the mask of training view ``index`` (NP/run_nerf_view.py:1049-1050; relative to the working
    directory there, to ``root`` here)."""
    return os.path.join(root, "Softmask", dataset_type, scene, "iter_500", f"softmask_{index:04d}_{top_k}per.png")


def write_softmask(root: str, dataset_type: str, scene: str, index: int, mask: np.ndarray, top_k: int = 30) -> str:
    """An 8-bit single-channel PNG; the script keeps ``value > 0`` as the mask (NP/run_nerf_view.py:1051,1054)."""
    path = softmask_path(root, dataset_type, scene, index, top_k)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    m = np.asarray(mask)
    _write_png(path, (m.astype(np.float32) * 255.0 + 0.5).astype(np.uint8) if m.dtype != np.uint8 else m)
    return path


def read_softmask(root: str, dataset_type: str, scene: str, index: int, top_k: int = 30) -> np.ndarray:
    """The boolean [H, W] mask the script builds from the file: (png / 255) > 0."""
    return (_read_png(softmask_path(root, dataset_type, scene, index, top_k)).astype(np.float32) / 255.0) > 0
