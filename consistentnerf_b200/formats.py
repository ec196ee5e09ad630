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

__all__ = ["save_checkpoint", "load_checkpoint", "latest_checkpoint", "read_pfm", "write_pfm", "write_metrics", "read_metrics"]

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
        out[k.strip()] = float(re.sub(r"[^0-9eE+\-.]", "", v.replace("tensor", "")) or "nan")
    return out
