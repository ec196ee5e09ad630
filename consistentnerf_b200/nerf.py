"""The reference's model-side surface (NP/run_nerf_helpers.py) on the CUDA library.

Same names, constructor arguments, state-dict keys and call conventions as the reference:
``NeRF``, ``Embedder``, ``get_embedder``, ``sample_pdf``, ``get_rays``, ``get_rays_np``, ``ndc_rays``,
``img2mse``, ``mse2psnr``, ``to8b``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops

__all__ = ["NeRF", "Embedder", "get_embedder", "sample_pdf", "get_rays", "get_rays_np", "ndc_rays",
           "img2mse", "mse2psnr", "to8b"]


# ---- misc (NP/run_nerf_helpers.py:9-11) ----------------------------------------------------
def img2mse(x, y):
    """mean((x-y)^2) through the K7 reduction kernel (differentiable)."""
    loss, stats = ops.MaskedMSEFn.apply(x, y, None, 1.0, 0.0, float(x.reshape(x.shape[0], -1).shape[0]), False)
    return loss


def mse2psnr(x):
    return -10.0 * torch.log(x) / torch.log(torch.tensor([10.0], device=x.device, dtype=x.dtype))


def to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


# ---- positional encoding (NP/run_nerf_helpers.py:15-63) --------------------------------------
class Embedder:
    """Callable object standing in for the reference's ``embed`` lambda.  It carries the
    octave count so ``run_network`` can recognise it and fuse the encoding into the MLP kernel."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs.get("input_dims", 3)
        if not kwargs.get("include_input", True) or not kwargs.get("log_sampling", True):
            raise NotImplementedError("only include_input=True, log_sampling=True (the reference's settings) are built")
        self.input_dims = d
        self.num_freqs = int(kwargs["num_freqs"])
        if int(kwargs["max_freq_log2"]) != self.num_freqs - 1:
            raise NotImplementedError("max_freq_log2 must equal num_freqs-1 (power-of-two octaves)")
        self.out_dim = d * (1 + 2 * self.num_freqs)

    def embed(self, inputs: torch.Tensor) -> torch.Tensor:
        if inputs.requires_grad:
            raise NotImplementedError("gradients w.r.t. encoded coordinates are not part of the hot path")
        shape = inputs.shape
        out = ops.posenc(inputs.reshape(-1, shape[-1]), self.num_freqs)
        return out.reshape(*shape[:-1], self.out_dim)

    __call__ = embed


class _IdentityEmbedder(nn.Identity):
    num_freqs = -1
    out_dim = 3


def get_embedder(multires, i=0):
    if i == -1:
        return _IdentityEmbedder(), 3
    e = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                 log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return e, e.out_dim


# ---- model (NP/run_nerf_helpers.py:67-130) --------------------------------------------------
class NeRF(nn.Module):
    """Parameter container identical to the reference module (same registration order, so the same
    seed gives the same initial weights and checkpoints are interchangeable).  ``forward`` runs the
    generic fp32 layer kernels; the renderer bypasses it with the fused tcgen05 kernel when the
    architecture is the canonical 8x256 one (see render.run_network)."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False,
                 coarse=False, stable_init=False):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips, self.use_viewdirs, self.coarse = skips, use_viewdirs, coarse
        self.temp_rgb = nn.Parameter(torch.full((1,), -0.7))
        self.temp_depth = nn.Parameter(torch.full((1,), -0.7))
        self.depth_scale = nn.Parameter(torch.full((1,), 1.0))
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        if stable_init:
            nn.init.uniform_(self.alpha_linear.bias)
        self.spec = ops.MLPSpec(D, W, input_ch, input_ch_views, output_ch, skips, use_viewdirs)
        self._packed = None

    def hot_params(self):
        """Tensors the kernels read, in MLPSpec.param_names() order (live storage, never copies)."""
        sd = dict(self.named_parameters())
        return [sd[n] for n in self.spec.param_names()]

    def packed_weights(self) -> "ops.PackedWeights":
        if self._packed is None:
            self._packed = ops.PackedWeights()
        return self._packed

    def forward(self, x):
        lead = x.shape[:-1]
        out = ops.LayerwiseMLPFn.apply(self.spec, x.reshape(-1, x.shape[-1]), *self.hot_params())
        return out.reshape(*lead, out.shape[-1])

    def load_weights_from_keras(self, weights):
        raise NotImplementedError("Keras weight import (NP/run_nerf_helpers.py:132) is outside the hot path")


# ---- ray helpers (NP/run_nerf_helpers.py:164-202) ---------------------------------------------
def get_rays(H, W, K, c2w):
    """[H,W,3] origins and directions of every pixel (generated on the GPU by cnerf_image_rays)."""
    dev = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else torch.device("cuda")
    rays = ops.image_rays(int(H), int(W), K, c2w, 0.0, 1.0, False, False, dev)
    return rays[:, 0:3].reshape(H, W, 3), rays[:, 3:6].reshape(H, W, 3)


def get_rays_np(H, W, K, c2w):
    """Host-side numpy twin used by the data samplers (NP/run_nerf_helpers.py:176-183)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    if float(near) != 1.0:
        raise NotImplementedError("the reference only ever calls ndc_rays with near=1 (NP/run_nerf.py:116)")
    shape = rays_d.shape
    rays = ops.pack_rays(rays_o, rays_d, 0.0, 1.0, False, ndc=(H, W, focal))
    return rays[:, 0:3].reshape(shape), rays[:, 3:6].reshape(shape)


# ---- hierarchical sampling (NP/run_nerf_helpers.py:206-250) -------------------------------------
def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    if bins.numel() == 0:                     # empty ray batch: the reference's shape, no kernel launch
        return torch.zeros(list(bins.shape[:-1]) + [N_samples], device=bins.device, dtype=torch.float32)
    u = None
    if pytest:
        np.random.seed(0)
        if not det:
            u = torch.tensor(np.random.rand(*(list(bins.shape[:-1]) + [N_samples])), dtype=torch.float32,
                             device=bins.device)
        else:
            u = torch.tensor(np.broadcast_to(np.linspace(0.0, 1.0, N_samples), list(bins.shape[:-1]) + [N_samples]).copy(),
                             dtype=torch.float32, device=bins.device)
    elif not det:
        u = torch.rand(list(bins.shape[:-1]) + [N_samples], device=bins.device)
    lead = bins.shape[:-1]
    out = ops.sample_pdf(bins.reshape(-1, bins.shape[-1]), weights.reshape(-1, weights.shape[-1]),
                         u.reshape(-1, N_samples) if u is not None else None, N_samples)
    return out.reshape(*lead, N_samples)
