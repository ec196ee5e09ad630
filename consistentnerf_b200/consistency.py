"""Cross-view consistency: depth-warp correspondence (K6) and masked losses (K7).

Reference surface kept: ``get_rays_ref``, ``get_ref_rays``, ``get_test_label`` (NP/run_nerf_view.py:553-669)
with their batched [B=1, ...] tensor conventions, so the unmodified hard-mask loop of ``train()``
(:999-1046) can call them.  ``build_hard_masks`` is the same precompute as one kernel per
(target, reference) pair, and ``masked_img_loss`` / ``masked_depth_loss`` are the loss expressions of
:1645-1648 / :1737 (and the cal_correspondance variants :1516-1517, :1550-1551) as single fused,
deterministic reductions without boolean-index compaction.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops

__all__ = ["get_rays_ref", "get_ref_rays", "get_test_label", "build_hard_masks", "masked_img_loss",
           "masked_depth_loss"]


def get_rays_ref(directions, c2w):
    """NP/run_nerf_view.py:553-574 (host-side tiny matmul; kept for API completeness)."""
    rays_d = directions @ c2w[:3, :3].T
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def _single_view(t, what):
    if t.shape[0] != 1:
        raise NotImplementedError(f"{what}: the reference only ever passes a batch of one view (B=1)")
    return t[0]


def get_ref_rays(w2c_ref, c2w_ref, intrinsic_ref, point_samples, img, depths_h=None):
    """NP/run_nerf_view.py:576-627.  Shapes as in the reference:
    returns rgb_ref [1,C,n_in], (depth_h_ref [1,1,n_in]), point_samples_cam [1,R,3], rays_o/rays_d [n_in,3],
    mask [1,R] bool, where n_in is the number of in-bounds projections."""
    B = point_samples.shape[0]
    pts = _single_view(point_samples.reshape(B, -1, 3), "get_ref_rays")
    N, C, H, W = img.shape
    res = ops.project_gather(pts, _single_view(w2c_ref, "get_ref_rays"), _single_view(intrinsic_ref, "get_ref_rays"),
                             H, W, img=_single_view(img, "get_ref_rays"),
                             depth=_single_view(depths_h, "get_ref_rays") if depths_h is not None else None,
                             c2w=_single_view(c2w_ref, "get_ref_rays"))
    mask = res["mask"].bool()
    rgb_ref = res["rgb"][mask].t().unsqueeze(0)                  # img[:, :, y[mask], x[mask]]
    rays_o, rays_d = res["rays_o"][mask], res["rays_d"][mask]
    cam = res["cam"].unsqueeze(0)
    if depths_h is not None:
        depth_ref = res["depth"][mask].reshape(1, 1, -1)
        return rgb_ref, depth_ref, cam, rays_o, rays_d, mask.unsqueeze(0)
    return rgb_ref, cam, rays_o, rays_d, mask.unsqueeze(0)


def get_test_label(w2c_ref, c2w_ref, intrinsic_ref, point_samples, img):
    """NP/run_nerf_view.py:630-669 -> (pixel_y [1,R], pixel_x [1,R], mask [1,R] bool, z_cam [1,R])."""
    B = point_samples.shape[0]
    pts = _single_view(point_samples.reshape(B, -1, 3), "get_test_label")
    N, C, H, W = img.shape
    res = ops.project_gather(pts, _single_view(w2c_ref, "get_test_label"),
                             _single_view(intrinsic_ref, "get_test_label"), H, W)
    return (res["py"].unsqueeze(0), res["px"].unsqueeze(0), res["mask"].bool().unsqueeze(0),
            res["cam"][:, 2].unsqueeze(0))


def build_hard_masks(rays_o_views: Sequence[torch.Tensor], rays_d_views: Sequence[torch.Tensor],
                     depths: Sequence[torch.Tensor], w2c_views: Sequence[torch.Tensor], K,
                     train_ids: Sequence[int], occlusion_threshold: float = 0.1, chunk: int = 5120,
                     rank: int = 0, world_size: int = 1):
    """The whole hard-mask precompute of train() (NP/run_nerf_view.py:999-1046) for the training views.

    ``rays_*_views[i]`` [H*W,3], ``depths[i]`` [H,W] prior depth and ``w2c_views[i]`` [4,4] are indexed by
    view id; returns {view id: bool mask [H,W]}.  ``chunk`` keeps the reference's per-5120-pixel
    threshold-doubling semantics.  With world_size > 1 the (target, reference) pairs are dealt round-robin
    to the ranks and the caller ORs the uint8 masks with one all-reduce (distributed.allreduce_masks)."""
    masks = {}
    pair = 0
    for tgt in train_ids:
        H, W = depths[tgt].shape
        m = torch.zeros(H * W, device=depths[tgt].device, dtype=torch.uint8)
        for ref in train_ids:
            if ref == tgt:
                continue
            if pair % world_size == rank:
                ops.hard_mask_pair(rays_o_views[tgt], rays_d_views[tgt], depths[tgt].reshape(-1), w2c_views[ref], K,
                                   depths[ref], thr0=occlusion_threshold, chunk=chunk, mask=m)
            pair += 1
        masks[tgt] = m.reshape(H, W)
    if world_size == 1:
        return {k: v.bool() for k, v in masks.items()}
    return masks


def masked_img_loss(rgb, target, mask, hardmask_coef: float, n_rand: Optional[int] = None, return_stats: bool = False,
                    global_counts=None):
    """img2mse(rgb[mask==1], target[mask==1]) + [mask.sum() != N_rand] * coef * img2mse(... mask==0).
    ``global_counts`` (distributed.global_mask_counts): the rows are one rank's shard of a global batch; the means run over the
    global batch and the returned value is this rank's additive share of the single-GPU loss (``n_rand`` = global N_rand).
    ``return_stats``: also the kernel's 5 device scalars [loss, #mask==1, #mask==0, plain img2mse over all rows, sum(mask)] --
    the quantities train() logs every step (NP/run_nerf_view.py:1908-1924) come out of the same launch, see ``loss_scalars``."""
    n_ref = float(rgb.shape[0] if n_rand is None else n_rand)
    if global_counts is not None and n_rand is None:
        raise ValueError("masked_img_loss: pass the GLOBAL n_rand together with global_counts")
    loss, stats = ops.MaskedMSEFn.apply(rgb, target, mask, 1.0, float(hardmask_coef), n_ref, True, global_counts)
    return (loss, stats) if return_stats else loss


def loss_scalars(stats: torch.Tensor, prefix: str = "") -> dict:
    """Named 0-d device tensors from a K7 stats vector: masked loss and its PSNR, plain MSE and its PSNR (mse2psnr,
    NP/run_nerf_helpers.py:11), mask population.  No host synchronisation: feed them to pipeline.StepLog."""
    s = stats.detach()
    to_psnr = lambda x: -10.0 * torch.log(x) / 2.302585092994046
    return {prefix + "masked_loss": s[0], prefix + "masked_psnr": to_psnr(s[0]), prefix + "mse": s[3], prefix + "psnr": to_psnr(s[3]),
            prefix + "n_masked": s[1]}


def masked_depth_loss(depth_pred, depth_prior, mask, far: float, hardmask_coef: float = 0.0,
                      n_rand: Optional[int] = None, include_unmasked: bool = False, return_stats: bool = False,
                      global_counts=None):
    """img2mse(d[mask==1]/far, prior[mask==1]/far) (+ coef * unmasked term in the cal_correspondance recipe)."""
    n_ref = float(depth_pred.shape[0] if n_rand is None else n_rand)
    loss, stats = ops.MaskedMSEFn.apply(depth_pred.reshape(-1, 1), depth_prior.reshape(-1, 1), mask, float(far),
                                        float(hardmask_coef), n_ref, bool(include_unmasked), global_counts)
    return (loss, stats) if return_stats else loss


def img2mse_softmask(x, y, temp):
    """sum(exp((x - y)^2 / temp) (x - y)^2) / sum(exp((x - y).detach()^2 / temp))   (NP/run_nerf_view.py:50); ``temp`` a float or the
    one-element device tensor softplus(network_fine.temp_rgb) (:1659), which then receives its gradient."""
    return ops.SoftMSEFn.apply(x, y, temp, 1.0, 0)


def img2mse_depth_softmask(x, y, temp):
    """The depth twin of img2mse_softmask (NP/run_nerf_view.py:55; same expression, called on depth / far, :1759)."""
    return ops.SoftMSEFn.apply(x, y, temp, 1.0, 0)


def img2mse_softLpmask(x, y, coef):
    """sum((|x - y|^coef + 1) (x - y)^2) / sum(|x - y|^coef + 1).detach()   (NP/run_nerf_view.py:58, --softLpmask, :1663-1664)."""
    return ops.SoftMSEFn.apply(x, y, float(coef), 1.0, 1)


def soft_depth_loss(depth_pred, depth_prior, far: float, param, kind: int = 1):
    """The scripts' depth form: the loss of depth_pred / far against depth_prior / far (NP/run_nerf_view.py:1759,1761) with the
    division done inside the kernel."""
    return ops.SoftMSEFn.apply(depth_pred, depth_prior, param if isinstance(param, torch.Tensor) else float(param), float(far), int(kind))
