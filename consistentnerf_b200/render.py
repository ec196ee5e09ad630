"""The reference's renderer-side surface on the CUDA library: ``render``, ``batchify_rays``,
``render_rays``, ``raw2outputs``, ``run_network``, ``batchify`` with the reference signatures.

Two flavours exist in the reference and both are served from here:

* ``run_nerf.py``                                  -> 3 maps + extras      (NP/run_nerf.py:70-137, :311-421)
* ``run_nerf_view.py`` / ``..._cal_correspondance.py`` -> + ``depth_map`` / ``depth0`` (NP/run_nerf_view.py:183-249, :441-551)

``make_api(with_depth)`` returns a namespace holding one flavour; the module-level names are the
``with_depth=True`` (ConsistentNeRF) flavour.  The call chain is kept exactly as in the reference
(render -> batchify_rays -> render_rays -> network_query_fn -> run_network), so any of these functions
can be patched individually into the unmodified scripts (see dropin.py).
"""
from __future__ import annotations

import types

import numpy as np
import torch

from . import ops
from .nerf import Embedder, NeRF

__all__ = ["batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays", "make_api"]

DEBUG = False


def batchify(fn, chunk):
    """NP/run_nerf.py:27-34."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def _fusable(fn, embed_fn, embeddirs_fn, inputs, viewdirs) -> bool:
    return (isinstance(fn, NeRF) and fn.spec.is_canonical and isinstance(embed_fn, Embedder)
            and isinstance(embeddirs_fn, Embedder) and embed_fn.num_freqs == 10 and embeddirs_fn.num_freqs == 4
            and viewdirs is not None and inputs.dim() == 3 and inputs.is_cuda
            and all(p.is_cuda for p in fn.hot_params()))


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """NP/run_nerf.py:37-52.  inputs [N_rays, N_samples, 3], viewdirs [N_rays, 3] -> [N_rays, N_samples, 4].

    Canonical network + reference encoders: one fused tcgen05 kernel (encoding never touches HBM,
    ``netchunk`` is irrelevant).  Anything else: encoding kernel + generic layer kernels."""
    if inputs.shape[0] == 0 and isinstance(fn, NeRF):      # empty ray batch: [0, S, 4|5] attached to the parameters, no kernel launch
        link = sum((p.reshape(-1)[:0].sum() for p in fn.parameters() if p.requires_grad), torch.zeros((), device=inputs.device))
        n_out = 4 if fn.use_viewdirs else fn.spec.output_ch
        return torch.zeros(list(inputs.shape[:-1]) + [n_out], device=inputs.device) + link
    if _fusable(fn, embed_fn, embeddirs_fn, inputs, viewdirs):
        params = fn.hot_params()
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in params)):      # inference: nothing to record
            packed = fn.packed_weights()
            packed.refresh(dict(zip(fn.spec.param_names(), params)))
            return ops.fused_mlp_forward(packed, inputs.detach(), viewdirs.detach())
        return ops.FusedMLPFn.apply(fn.spec, fn.packed_weights(), embed_fn.num_freqs, embeddirs_fn.num_freqs,
                                    inputs, viewdirs, *params)
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded = torch.cat([embedded, embeddirs_fn(input_dirs_flat)], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """NP/run_nerf.py:265-308 -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    if raw.shape[0] == 0:                     # empty ray batch: the reference's shapes, attached to `raw`, no kernel launch
        total = raw[..., 3].sum(-1)
        return raw[..., :3].sum(-2), total, total, raw[..., 3], total
    noise = None
    if raw_noise_std > 0.0:
        if pytest:      # fixed-RNG hook of the reference (:290-294): *uniform* numbers from numpy seed 0
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std, dtype=torch.float32,
                                 device=raw.device)
        else:
            noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
    return ops.CompositeFn.apply(raw, z_vals, rays_d, noise, bool(white_bkgd))


def _render_rays(with_depth, ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False,
                 perturb=0.0, N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.0, verbose=False,
                 pytest=False):
    if not ray_batch.is_cuda:
        raise RuntimeError("render_rays: ray_batch must be a CUDA tensor (no CPU path)")
    dev = ray_batch.device
    rays = ray_batch if (ray_batch.dtype == torch.float32 and ray_batch.is_contiguous()) else ray_batch.float().contiguous()
    N_rays = rays.shape[0]
    if N_rays == 0:
        return _empty_result(with_depth, dev, network_fn, network_fine, N_samples, N_importance, retraw)
    rays_d = rays[:, 3:6]
    viewdirs = rays[:, -3:].contiguous() if rays.shape[-1] > 8 else None

    t_vals = ops.unit_linspace(N_samples, dev)
    t_rand = None
    if perturb > 0.0:
        if pytest:      # NP/run_nerf.py:376-380
            np.random.seed(0)
            t_rand = torch.tensor(np.random.rand(N_rays, N_samples), dtype=torch.float32, device=dev)
        else:
            t_rand = torch.rand((N_rays, N_samples), device=dev)
    z_vals, pts = ops.stratified(rays, t_vals, t_rand, lindisp)                       # K1

    raw = network_query_fn(pts, viewdirs, network_fn)                                 # K2+K3
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd,
                                                                 pytest=pytest)      # K4
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0, depth_map_0 = rgb_map, disp_map, acc_map, depth_map
        u = None
        if pytest:      # NP/run_nerf_helpers.py:220-229
            np.random.seed(0)
            if perturb != 0.0:
                u = torch.tensor(np.random.rand(N_rays, N_importance), dtype=torch.float32, device=dev)
        elif perturb != 0.0:
            u = torch.rand((N_rays, N_importance), device=dev)
        z_samples, z_vals, z_std = ops.sample_fine(z_vals, weights.detach(), u, N_importance)   # K5 (+sort, std)
        pts = ops.ray_points(rays, z_vals)
        run_fn = network_fn if network_fine is None else network_fine
        raw = network_query_fn(pts, viewdirs, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd,
                                                                     pytest=pytest)

    ret = {"rgb_map": rgb_map, "disp_map": disp_map, "acc_map": acc_map}
    if with_depth:
        ret["depth_map"] = depth_map
    if retraw:
        ret["raw"] = raw
    if N_importance > 0:
        ret["rgb0"], ret["disp0"], ret["acc0"] = rgb_map_0, disp_map_0, acc_map_0
        if with_depth:
            ret["depth0"] = depth_map_0
        ret["z_std"] = z_std
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


def _empty_result(with_depth, dev, network_fn, network_fine, N_samples, N_importance, retraw):
    """render_rays on an empty ray batch: the reference's shapes ([0, ...]) with no kernel launch (a zero-sized grid is an
    error); the tensors stay attached to the parameters so that loss.backward() yields zero gradients, as nn.Linear does
    on an empty input."""
    params = [p for net in (network_fn, network_fine) if isinstance(net, torch.nn.Module) for p in net.parameters() if p.requires_grad]
    link = sum((p.reshape(-1)[:0].sum() for p in params), torch.zeros((), device=dev)) if torch.is_grad_enabled() else torch.zeros((), device=dev)

    def z(*shape):
        return torch.zeros(shape, device=dev) + link
    S = N_samples + (N_importance if N_importance > 0 else 0)
    ret = {"rgb_map": z(0, 3), "disp_map": z(0), "acc_map": z(0)}
    if with_depth:
        ret["depth_map"] = z(0)
    if retraw:
        ret["raw"] = z(0, S, 4)
    if N_importance > 0:
        ret["rgb0"], ret["disp0"], ret["acc0"] = z(0, 3), z(0), z(0)
        if with_depth:
            ret["depth0"] = z(0)
        ret["z_std"] = torch.zeros((0,), device=dev)
    return ret


def make_api(with_depth: bool) -> types.SimpleNamespace:
    """Namespace with render / batchify_rays / render_rays of one reference flavour."""
    api = types.SimpleNamespace(batchify=batchify, run_network=run_network, raw2outputs=raw2outputs)

    def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.0,
                    N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.0, verbose=False, pytest=False):
        return _render_rays(with_depth, ray_batch, network_fn, network_query_fn, N_samples, retraw, lindisp, perturb,
                            N_importance, network_fine, white_bkgd, raw_noise_std, verbose, pytest)

    def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
        """NP/run_nerf.py:55-67.  Results are chunk invariant (every kernel is per-ray)."""
        all_ret = {}
        if rays_flat.shape[0] == 0:
            return api.render_rays(rays_flat, **kwargs)
        for i in range(0, rays_flat.shape[0], chunk):
            ret = api.render_rays(rays_flat[i:i + chunk], **kwargs)
            for k in ret:
                all_ret.setdefault(k, []).append(ret[k])
        return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}

    def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0.0, far=1.0, use_viewdirs=False,
               c2w_staticcam=None, **kwargs):
        """NP/run_nerf.py:70-137 / NP/run_nerf_view.py:183-249."""
        if c2w is not None and c2w_staticcam is None:
            # whole image: rays are generated, normalised, NDC-warped and packed by one kernel
            dev = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else torch.device("cuda")
            packed = ops.image_rays(int(H), int(W), K, c2w, near, far, use_viewdirs, bool(ndc), dev)
            sh = (int(H), int(W), 3)
        else:
            if c2w is not None:      # c2w_staticcam: view directions from c2w, geometry from the static camera
                from .nerf import get_rays
                _, viewdirs_src = get_rays(H, W, K, c2w)
                rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
            else:
                rays_o, rays_d = rays
                viewdirs_src = None
            if not rays_d.is_cuda:
                raise RuntimeError("render: rays must be CUDA tensors (no CPU path)")
            sh = tuple(rays_d.shape)
            packed = ops.pack_rays(rays_o, rays_d, near, far, use_viewdirs, ndc=(H, W, K[0][0]) if ndc else None)
            if viewdirs_src is not None and use_viewdirs:
                v = viewdirs_src.reshape(-1, 3)
                packed[:, 8:11] = v / torch.norm(v, dim=-1, keepdim=True)
        all_ret = api.batchify_rays(packed, chunk, **kwargs)
        for k in all_ret:
            all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
        k_extract = ["rgb_map", "disp_map", "acc_map"] + (["depth_map"] if with_depth else [])
        ret_list = [all_ret[k] for k in k_extract]
        ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
        return ret_list + [ret_dict]

    api.render_rays, api.batchify_rays, api.render = render_rays, batchify_rays, render
    return api


_view = make_api(with_depth=True)
render_rays, batchify_rays, render = _view.render_rays, _view.batchify_rays, _view.render
