"""The callers on either side of the hot path, re-designed for a GPU-resident pipeline.

* ``RayBank``     -- the per-step batch sampler of train() (NP/run_nerf_view.py:1443-1517, ``--no_batching``; NP/run_nerf.py:
  718-760).  The reference re-uploads the whole target image / prior depth / mask of the chosen view and runs get_rays over
  the full image every step, then gathers ``N_rand`` pixels with numpy indices.  Here every view lives on the device once;
  a step draws the view on the host (one integer), the pixel ids on the device, and one kernel emits the packed rays and
  the gathered targets of exactly those pixels.
* ``StepLog``     -- the per-step scalar logging of train() (NP/run_nerf_view.py:1908-1937): ten-odd ``add_scalar(tensor)`` calls per
  step, each a device synchronisation.  Here the scalars stay on the device in a ring of rows and reach the host with one
  copy every ``i_print`` steps.
* ``render_path`` -- the novel-view image loop (NP/run_nerf_view.py:252-294): images are rendered back to back while the
  previous image's device->host copy runs on a side stream into pinned double buffers; same return value
  ``(rgbs, disps, accs)`` as the reference.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .nerf import to8b

__all__ = ["RayBank", "render_path", "StepLog"]


class RayBank:
    """Device-resident training views + random ray batches."""

    def __init__(self, images, poses, K, near: float, far: float, use_viewdirs: bool = True, ndc: bool = False,
                 depths=None, masks=None, device="cuda", seed: int = 0):
        dev = torch.device(device)
        self.images = torch.as_tensor(np.asarray(images), dtype=torch.float32).to(dev).contiguous()     # [V,H,W,3]
        self.V, self.H, self.W = self.images.shape[:3]
        self.poses = torch.as_tensor(np.asarray(poses), dtype=torch.float32)[:, :3, :4].contiguous()    # host: tiny
        self.K = [[float(K[i][j]) for j in range(3)] for i in range(3)]
        self.depths = torch.as_tensor(np.asarray(depths), dtype=torch.float32).to(dev).contiguous() if depths is not None else None
        self.masks = torch.as_tensor(np.asarray(masks), dtype=torch.float32).to(dev).contiguous() if masks is not None else None
        self.near, self.far, self.use_viewdirs, self.ndc = float(near), float(far), bool(use_viewdirs), bool(ndc)
        self.device = dev
        self.host_rng = np.random.RandomState(seed)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)

    def sample(self, n_rand: int, view: Optional[int] = None, views: Optional[Sequence[int]] = None,
               precrop_frac: Optional[float] = None, patches: int = 0, patch_size: int = 16):
        """One ``--no_batching`` batch: ``patches`` random ``patch_size``^2 patches (the MiDaS / SSIM patches of
        NP/run_nerf_view.py:1472-1503, without its white-background rejection) followed by ``n_rand`` distinct random
        pixels, optionally restricted to the centre crop (:1455-1463).  Returns a dict with ``rays`` [n,8|11] (packed like
        render()'s ray batch), ``batch_rays`` [2,n,3], ``target`` [n,3], ``depth`` [n], ``mask`` [n], ``pix`` [n], ``view``."""
        H, W, dev = self.H, self.W, self.device
        if view is None:
            pool = list(range(self.V)) if views is None else list(views)
            view = int(pool[self.host_rng.randint(len(pool))])
        if precrop_frac is not None:
            dH, dW = int(H // 2 * precrop_frac), int(W // 2 * precrop_frac)
            y0, x0, h, w = H // 2 - dH, W // 2 - dW, 2 * dH, 2 * dW
        else:
            y0, x0, h, w = 0, 0, H, W
        sel = torch.randperm(h * w, device=dev, generator=self.gen)[:n_rand]            # distinct pixels (replace=False)
        pix = (sel // w + y0) * W + (sel % w + x0)
        if patches > 0:
            py = torch.randint(y0, y0 + h - patch_size + 1, (patches,), device=dev, generator=self.gen)
            px = torch.randint(x0, x0 + w - patch_size + 1, (patches,), device=dev, generator=self.gen)
            ar = torch.arange(patch_size, device=dev)
            rows = (py[:, None, None] + ar[None, :, None]).expand(patches, patch_size, patch_size)
            cols = (px[:, None, None] + ar[None, None, :]).expand(patches, patch_size, patch_size)
            pix = torch.cat([(rows * W + cols).reshape(-1), pix])
        rays, target, depth, mask = ops.gather_rays(
            H, W, self.K, self.poses[view], pix, self.near, self.far, self.use_viewdirs, self.ndc, self.images[view],
            self.depths[view] if self.depths is not None else None, self.masks[view] if self.masks is not None else None)
        return {"rays": rays, "batch_rays": torch.stack([rays[:, 0:3], rays[:, 3:6]], 0), "target": target, "depth": depth,
                "mask": mask, "pix": pix, "view": view}


class StepLog:
    """Per-step training scalars without per-step host synchronisation.

        log = StepLog(["loss", "psnr", "psnr0"], capacity=100)
        log.record(i, loss=loss, **loss_scalars(stats))      # 0-d device tensors (or Python floats); stream ordered, no sync
        if i % i_print == 0:
            for step, row in log.flush():                    # ONE device->host copy for everything recorded since the last flush
                writer.add_scalar(...)
    """

    def __init__(self, names: Sequence[str], capacity: int = 1024, device="cuda"):
        self.names = list(names)
        self.col = {n: j for j, n in enumerate(self.names)}
        self.capacity = int(capacity)
        self.buf = torch.full((self.capacity, len(self.names)), float("nan"), device=device, dtype=torch.float32)
        self.steps = []

    def record(self, step: int, **scalars):
        if len(self.steps) >= self.capacity:
            raise RuntimeError("StepLog is full: call flush() at least every `capacity` steps")
        row = self.buf[len(self.steps)]
        for name, v in scalars.items():
            j = self.col.get(name)
            if j is None:
                continue                                    # not a logged quantity
            if isinstance(v, torch.Tensor):
                row[j:j + 1].copy_(v.detach().reshape(1), non_blocking=True)
            else:
                row[j] = float(v)
        self.steps.append(int(step))

    def flush(self):
        """[(step, {name: float})] of every record since the last flush; one device->host copy."""
        n = len(self.steps)
        if n == 0:
            return []
        host = self.buf[:n].cpu().numpy()
        out = [(st, {name: float(host[i, j]) for name, j in self.col.items()}) for i, st in enumerate(self.steps)]
        self.buf[:n].fill_(float("nan"))
        self.steps = []
        return out


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0, render_fn=None):
    """NP/run_nerf_view.py:252-294 -> (rgbs [V,H,W,3], disps [V,H,W], accs [V,H,W]) as numpy arrays.

    Image i+1 is rendered while image i travels device->host on a side stream into one of two pinned buffers; nothing
    synchronises the render stream inside the loop.  ``savedir`` writes ``color_%03d.png`` with cv2 when available."""
    if render_fn is None:
        from .render import render as render_fn
    H, W, focal = hwf
    H, W = int(H), int(W)
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    n = len(render_poses)
    dev = torch.device("cuda")
    rgbs = np.empty((n, H, W, 3), dtype=np.float32)
    disps = np.empty((n, H, W), dtype=np.float32)
    accs = np.empty((n, H, W), dtype=np.float32)
    pinned = [torch.empty((H, W, 5), dtype=torch.float32, device="cpu", pin_memory=True) for _ in range(2)]
    staged = [torch.empty((H, W, 5), dtype=torch.float32, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    done = [None, None]

    def flush(slot, idx):
        done[slot].synchronize()
        a = pinned[slot].numpy()
        rgbs[idx], disps[idx], accs[idx] = a[..., :3], a[..., 3], a[..., 4]
        if savedir is not None:
            try:
                import cv2
                cv2.imwrite(os.path.join(savedir, "color_{:03d}.png".format(idx)), to8b(rgbs[idx])[..., ::-1])
            except ImportError:
                pass

    with torch.no_grad():
        for i, c2w in enumerate(render_poses):
            slot = i & 1
            if i >= 2:
                flush(slot, i - 2)                      # buffer about to be reused
            c2w_t = c2w if isinstance(c2w, torch.Tensor) else torch.as_tensor(np.asarray(c2w), dtype=torch.float32)
            out = render_fn(H, W, K, chunk=chunk, c2w=c2w_t[:3, :4], **render_kwargs)
            rgb, disp, acc = out[0], out[1], out[2]
            staged[slot][..., :3].copy_(rgb)
            staged[slot][..., 3].copy_(disp)
            staged[slot][..., 4].copy_(acc)
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev)
                pinned[slot].copy_(staged[slot], non_blocking=True)
                done[slot] = torch.cuda.Event()
                done[slot].record()
        for i in range(max(0, n - 2), n):
            flush(i & 1, i)
    return rgbs, disps, accs
