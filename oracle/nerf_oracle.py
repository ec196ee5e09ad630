"""CPU restatement of the reference's per-ray volumetric rendering path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function is dtype generic
(float32 reproduces the reference arithmetic on CPU, float64 is the adjudicator of
SURVEY.md section 8c) and cites the reference lines it restates.  ``NP/`` below means
``/root/reference/nerf-pytorch-master/``.

The functions are written functionally (explicit parameter dictionaries, explicit
random tensors) so that a CUDA kernel and the oracle can be fed *identical* inputs;
the reference draws its random numbers inside the functions instead.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

__all__ = [
    "posenc", "nerf_param_shapes", "make_params", "mlp_forward", "stratified_z",
    "ray_points", "composite", "sample_pdf", "merge_sorted", "render_rays",
    "pixel_rays", "ndc_warp", "pack_rays", "project_points", "gather_reference",
    "reference_view_rays", "hard_mask_pair", "masked_mse", "mse_to_psnr", "aten_cpu_sum_f32",
]


# --------------------------------------------------------------------------------------
# positional encoding  (NP/run_nerf_helpers.py:15-46 Embedder, :48-63 get_embedder)
# --------------------------------------------------------------------------------------
def posenc(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """[..., C] -> [..., C*(1+2*n_freqs)]: input first, then for every octave
    sin(2^k x) followed by cos(2^k x), each C wide (NP/run_nerf_helpers.py:24-46).
    ``n_freqs < 0`` is the identity embedder (i_embed == -1, :49-50)."""
    if n_freqs < 0:
        return x
    pieces = [x]
    for k in range(n_freqs):
        # freq_bands = 2 ** linspace(0, L-1, L): exact powers of two (:32)
        f = float(2.0 ** k)
        pieces.append(torch.sin(x * f))
        pieces.append(torch.cos(x * f))
    return torch.cat(pieces, dim=-1)


# --------------------------------------------------------------------------------------
# NeRF MLP  (NP/run_nerf_helpers.py:67-130)
# --------------------------------------------------------------------------------------
def nerf_param_shapes(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5,
                      skips=(4,), use_viewdirs=True) -> List[Tuple[str, Tuple[int, ...]]]:
    """State-dict keys and shapes in the module's registration order
    (NP/run_nerf_helpers.py:78-101)."""
    out = [("temp_rgb", (1,)), ("temp_depth", (1,)), ("depth_scale", (1,))]
    fan_in = input_ch
    for i in range(D):
        out.append((f"pts_linears.{i}.weight", (W, fan_in)))
        out.append((f"pts_linears.{i}.bias", (W,)))
        # layer i+1 sees the re-concatenated input when i is a skip index (:86-87,113-114)
        fan_in = W + input_ch if i in skips else W
    out.append(("views_linears.0.weight", (W // 2, input_ch_views + W)))
    out.append(("views_linears.0.bias", (W // 2,)))
    if use_viewdirs:
        out += [("feature_linear.weight", (W, W)), ("feature_linear.bias", (W,)),
                ("alpha_linear.weight", (1, W)), ("alpha_linear.bias", (1,)),
                ("rgb_linear.weight", (3, W // 2)), ("rgb_linear.bias", (3,))]
    else:
        out += [("output_linear.weight", (output_ch, W)), ("output_linear.bias", (output_ch,))]
    return out


def make_params(seed: int, dtype=torch.float32, sigma_bias: float = 0.0, **arch) -> Dict[str, torch.Tensor]:
    """Deterministic parameters drawn with numpy's RandomState (stable across torch
    versions) at the scale of nn.Linear's default init (uniform +-1/sqrt(fan_in)).
    ``sigma_bias`` shifts the density head so that a chosen share of samples is
    occupied ("trained-like" set of SURVEY.md section 8d)."""
    rs = np.random.RandomState(seed)
    params = {}
    for name, shape in nerf_param_shapes(**arch):
        if name in ("temp_rgb", "temp_depth"):
            v = np.full(shape, -0.7)
        elif name == "depth_scale":
            v = np.full(shape, 1.0)
        else:
            layer = name.rsplit(".", 1)[0]
            fan_in = dict(nerf_param_shapes(**arch))[layer + ".weight"][1]
            bound = 1.0 / math.sqrt(fan_in)
            v = rs.uniform(-bound, bound, size=shape)
        params[name] = torch.from_numpy(np.asarray(v, dtype=np.float64)).to(dtype)
    if sigma_bias:
        key = "alpha_linear.bias" if "alpha_linear.bias" in params else "output_linear.bias"
        params[key] = params[key].clone()
        params[key][0 if key.startswith("alpha") else 3] += sigma_bias   # density channel
    return params


def _affine(params, name, h):
    return h @ params[name + ".weight"].t() + params[name + ".bias"]


def mlp_forward(params: Dict[str, torch.Tensor], x: torch.Tensor, D=8, W=256, input_ch=63,
                input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True) -> torch.Tensor:
    """x [P, input_ch+input_ch_views] -> [P, 4] (viewdirs) or [P, output_ch]
    (NP/run_nerf_helpers.py:107-130)."""
    pts, views = x[..., :input_ch], x[..., input_ch:input_ch + input_ch_views]
    h = pts
    for i in range(D):
        h = torch.relu(_affine(params, f"pts_linears.{i}", h))
        if i in skips:
            h = torch.cat([pts, h], dim=-1)          # input first (:113-114)
    if not use_viewdirs:
        return _affine(params, "output_linear", h)
    sigma = _affine(params, "alpha_linear", h)       # no activation (:117)
    feat = _affine(params, "feature_linear", h)      # no activation (:118)
    hv = torch.relu(_affine(params, "views_linears.0", torch.cat([feat, views], dim=-1)))
    rgb = _affine(params, "rgb_linear", hv)
    return torch.cat([rgb, sigma], dim=-1)


# --------------------------------------------------------------------------------------
# stratified sampling  (NP/run_nerf.py:360-384)
# --------------------------------------------------------------------------------------
def stratified_z(near: torch.Tensor, far: torch.Tensor, n_samples: int, lindisp: bool = False,
                 t_rand: Optional[torch.Tensor] = None, t_vals: Optional[torch.Tensor] = None) -> torch.Tensor:
    """near/far [N,1] -> z [N,S].  ``t_rand`` [N,S] in [0,1) switches the jitter on
    (perturb > 0 branch, :368-382); None keeps the bin edges (:361-366).  ``t_vals`` overrides
    the linspace (CPU and CUDA linspace may differ in the last bit)."""
    t = torch.linspace(0.0, 1.0, n_samples, dtype=near.dtype) if t_vals is None else t_vals.to(near.dtype)
    if lindisp:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    else:
        z = near * (1.0 - t) + far * t
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], dim=-1)
        lo = torch.cat([z[:, :1], mid], dim=-1)
        z = lo + (hi - lo) * t_rand
    return z


def ray_points(rays_o, rays_d, z):
    """o + d * z  -> [N,S,3]  (NP/run_nerf.py:384,400)."""
    return rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]


# --------------------------------------------------------------------------------------
# alpha compositing  (NP/run_nerf.py:265-308)
# --------------------------------------------------------------------------------------
def composite(raw: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor,
              noise: Optional[torch.Tensor] = None, white_bkgd: bool = False) -> Dict[str, torch.Tensor]:
    """raw [N,S,4], z [N,S], rays_d [N,3]; ``noise`` [N,S] is the *already scaled*
    density noise (randn * raw_noise_std, :286-288)."""
    delta = z[:, 1:] - z[:, :-1]
    delta = torch.cat([delta, torch.full_like(delta[:, :1], 1e10)], dim=-1)      # :280-281
    delta = delta * torch.norm(rays_d[:, None, :], dim=-1)                       # :283
    colour = torch.sigmoid(raw[..., :3])                                         # :285
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * delta)                          # :278,296
    ones = torch.ones((alpha.shape[0], 1), dtype=alpha.dtype)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], dim=-1), dim=-1)[:, :-1]
    w = alpha * trans                                                            # :298
    rgb = torch.sum(w[..., None] * colour, dim=-2)
    depth = torch.sum(w * z, dim=-1)
    acc = torch.sum(w, dim=-1)
    disp = 1.0 / torch.max(1e-10 * torch.ones_like(depth), depth / acc)          # :302
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])                                       # :305-306
    return {"rgb": rgb, "disp": disp, "acc": acc, "weights": w, "depth": depth}


# --------------------------------------------------------------------------------------
# hierarchical sampling  (NP/run_nerf_helpers.py:206-250)
# --------------------------------------------------------------------------------------
def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, u: torch.Tensor, return_debug: bool = False):
    """bins [N,B], weights [N,B-1], u [N,M] -> samples [N,M].  With ``return_debug``
    also the cdf and the below/above bin indices the reference keeps internal
    (:233-236) -- the quantities the "bit-exact indices" contract is stated on."""
    w = weights + 1e-5                                                           # :208
    pdf = w / torch.sum(w, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)               # :211
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)                                 # :233
    below = torch.clamp(idx - 1, min=0)
    above = torch.clamp(idx, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    span = cdf_a - cdf_b
    span = torch.where(span < 1e-5, torch.ones_like(span), span)                 # :245-246
    frac = (u - cdf_b) / span
    out = bin_b + frac * (bin_a - bin_b)
    if return_debug:
        return out, {"cdf": cdf, "below": below, "above": above}
    return out


def merge_sorted(z_coarse: torch.Tensor, z_new: torch.Tensor) -> torch.Tensor:
    """sort(cat) along the ray (NP/run_nerf.py:399)."""
    return torch.sort(torch.cat([z_coarse, z_new], dim=-1), dim=-1).values


# --------------------------------------------------------------------------------------
# render_rays  (NP/run_nerf.py:311-421; depth outputs as NP/run_nerf_view.py:441-551)
# --------------------------------------------------------------------------------------
def _query(params, arch, pts, viewdirs, multires, multires_views):
    """run_network (NP/run_nerf.py:37-52) without the chunking loops."""
    n, s = pts.shape[:2]
    e = posenc(pts.reshape(-1, 3), multires)
    if viewdirs is not None:
        d = viewdirs[:, None, :].expand(n, s, 3).reshape(-1, 3)
        e = torch.cat([e, posenc(d, multires_views)], dim=-1)
    return mlp_forward(params, e, **arch).reshape(n, s, -1)


def render_rays(ray_batch: torch.Tensor, coarse: Dict[str, torch.Tensor],
                fine: Optional[Dict[str, torch.Tensor]], arch: dict, n_samples: int,
                n_importance: int = 0, multires: int = 10, multires_views: int = 4,
                lindisp: bool = False, white_bkgd: bool = False,
                t_rand: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
                noise_coarse: Optional[torch.Tensor] = None, noise_fine: Optional[torch.Tensor] = None,
                retraw: bool = False) -> Dict[str, torch.Tensor]:
    """The whole per-ray pipeline.  ``t_rand`` None <=> perturb == 0 and then ``u`` None
    means the deterministic linspace of sample_pdf(det=True)."""
    dt = ray_batch.dtype
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z = stratified_z(near, far, n_samples, lindisp, t_rand)
    raw = _query(coarse, arch, ray_points(rays_o, rays_d, z), viewdirs, multires, multires_views)
    c = composite(raw, z, rays_d, noise_coarse, white_bkgd)
    out = {"rgb_map": c["rgb"], "disp_map": c["disp"], "acc_map": c["acc"], "depth_map": c["depth"],
           "z_coarse": z, "weights_coarse": c["weights"], "raw_coarse": raw}
    if n_importance > 0:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        if u is None:
            u = torch.linspace(0.0, 1.0, n_importance, dtype=dt).expand(z.shape[0], n_importance)
        z_new, dbg = sample_pdf(mid, c["weights"][:, 1:-1], u, return_debug=True)
        z_new = z_new.detach()
        z_all = merge_sorted(z, z_new)
        net = coarse if fine is None else fine
        raw_f = _query(net, arch, ray_points(rays_o, rays_d, z_all), viewdirs, multires, multires_views)
        f = composite(raw_f, z_all, rays_d, noise_fine, white_bkgd)
        out.update({"rgb0": c["rgb"], "disp0": c["disp"], "acc0": c["acc"], "depth0": c["depth"],
                    "rgb_map": f["rgb"], "disp_map": f["disp"], "acc_map": f["acc"],
                    "depth_map": f["depth"], "z_std": torch.std(z_new, dim=-1, unbiased=False),
                    "z_samples": z_new, "z_fine": z_all, "weights_fine": f["weights"],
                    "pdf_debug": dbg})
        raw = raw_f
    if retraw:
        out["raw"] = raw
    return out


# --------------------------------------------------------------------------------------
# ray preparation  (NP/run_nerf_helpers.py:164-202, NP/run_nerf.py:95-126)
# --------------------------------------------------------------------------------------
def pixel_rays(H: int, W: int, K, c2w: torch.Tensor):
    """Pinhole rays for every pixel, row-major [H,W,3] (NP/run_nerf_helpers.py:164-173)."""
    dt = c2w.dtype
    jj, ii = torch.meshgrid(torch.linspace(0, H - 1, H, dtype=dt), torch.linspace(0, W - 1, W, dtype=dt),
                            indexing="ij")
    dirs = torch.stack([(ii - K[0][2]) / K[0][0], -(jj - K[1][2]) / K[1][1], -torch.ones_like(ii)], dim=-1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], dim=-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_warp(H, W, focal, near, rays_o, rays_d):
    """NP/run_nerf_helpers.py:186-202."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1.0 / (W / (2.0 * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1.0 / (H / (2.0 * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1.0 + 2.0 * near / rays_o[..., 2]
    d0 = -1.0 / (W / (2.0 * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1.0 / (H / (2.0 * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2.0 * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], dim=-1), torch.stack([d0, d1, d2], dim=-1)


def pack_rays(rays_o, rays_d, near: float, far: float, use_viewdirs: bool, ndc=None):
    """[N,8|11] ray batch as assembled in render() (NP/run_nerf.py:101-126).
    ``ndc`` = (H, W, focal) applies the NDC warp after the view directions were taken."""
    vd = None
    if use_viewdirs:
        vd = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        vd = vd.reshape(-1, 3)
    if ndc is not None:
        rays_o, rays_d = ndc_warp(ndc[0], ndc[1], ndc[2], 1.0, rays_o, rays_d)
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    cols = [rays_o, rays_d, near * torch.ones_like(rays_d[:, :1]), far * torch.ones_like(rays_d[:, :1])]
    if vd is not None:
        cols.append(vd)
    return torch.cat(cols, dim=-1)


# --------------------------------------------------------------------------------------
# cross-view consistency geometry  (NP/run_nerf_view.py:553-669, :999-1046;
# CPU twin of the same code: RG/internal/mask_utils/mask_generator.py:82-133)
# --------------------------------------------------------------------------------------
def project_points(pts_w: torch.Tensor, w2c: torch.Tensor, K: torch.Tensor, H: int, W: int):
    """World points [R,3] -> reference-view pixel (x, y as rounded floats), the strict
    in-bounds mask and the camera-space point.  NP/run_nerf_view.py:594-613.

    The three-term dot products are evaluated left to right without fusing so that a
    kernel can reproduce them bit for bit (torch's matmul order is not specified)."""
    R, T = w2c[:3, :3], w2c[:3, 3]
    cam = []
    for r in range(3):
        v = pts_w[:, 0] * R[r, 0]
        v = v + pts_w[:, 1] * R[r, 1]
        v = v + pts_w[:, 2] * R[r, 2]
        cam.append(v + T[r])
    xc, yc, zc = cam[0], -cam[1], -cam[2]           # @ diag(1,-1,-1)  (:596-597)
    pix = []
    for r in range(3):
        v = xc * K[r, 0]
        v = v + yc * K[r, 1]
        v = v + zc * K[r, 2]
        pix.append(v)
    px = torch.round(pix[0] / pix[2] + 0.0)         # half-to-even (:604-605)
    py = torch.round(pix[1] / pix[2] + 0.0)
    nx, ny = px / (W - 1), py / (H - 1)             # :607-608
    mask = (nx > 0.0) & (nx < 1.0) & (ny > 0.0) & (ny < 1.0)      # strict both sides (:611-613)
    return px, py, mask, torch.stack([xc, yc, zc], dim=-1)


def gather_reference(img_chw: torch.Tensor, depth_hw: Optional[torch.Tensor], px, py, mask):
    """Integer gather of the reference image/depth at the masked pixels (:622-624).
    Returns dense [R,C] / [R] tensors that are zero where mask is False."""
    xi = torch.where(mask, px, torch.zeros_like(px)).long()
    yi = torch.where(mask, py, torch.zeros_like(py)).long()
    rgb = img_chw[:, yi, xi].t() * mask[:, None].to(img_chw.dtype)
    dep = None
    if depth_hw is not None:
        dep = depth_hw[yi, xi] * mask.to(depth_hw.dtype)
    return rgb, dep


def reference_view_rays(px, py, K: torch.Tensor, c2w: torch.Tensor):
    """Rays of the reference view through the rounded pixels (:615-620, get_rays_ref :553-574)."""
    dirs = torch.stack([(px - K[0, 2]) / K[0, 0], (py - K[1, 2]) / K[1, 1], torch.ones_like(px)], dim=-1)
    rays_d = dirs @ c2w[:3, :3].t()
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return rays_o, rays_d


def hard_mask_pair(rays_o, rays_d, depth_tgt, w2c_ref, K, ref_depth_hw, thr0: float = 0.1,
                   chunk: int = 5120) -> torch.Tensor:
    """Occlusion-aware correspondence mask of one (target, reference) view pair.
    NP/run_nerf_view.py:1014-1039: per ``chunk`` pixels, back-project with the prior
    depth, project into the reference view, keep pixels whose camera depth agrees with
    the reference prior depth within a threshold that doubles until the chunk has a hit."""
    H, W = ref_depth_hw.shape
    n = depth_tgt.shape[0]
    out = torch.zeros(n, dtype=torch.bool)
    for s in range(0, n, chunk):
        e = min(s + chunk, n)
        pw = rays_o[s:e] + depth_tgt[s:e, None] * rays_d[s:e]
        px, py, inb, cam = project_points(pw, w2c_ref, K, H, W)
        if int(inb.sum()) == 0:
            continue
        _, dref = gather_reference(ref_depth_hw[None], ref_depth_hw, px, py, inb)
        diff = (cam[:, 2] - dref).abs()
        thr = torch.tensor(thr0, dtype=diff.dtype)
        hit = inb & (diff < thr)
        guard = 0
        while int(hit.sum()) == 0 and guard < 200:
            thr = thr * 2
            hit = inb & (diff < thr)
            guard += 1
        out[s:e] = hit
    return out


# --------------------------------------------------------------------------------------
# masked consistency losses  (NP/run_nerf_view.py:1645-1648,1737; cal_correspondance :1516-1517,1550-1551)
# --------------------------------------------------------------------------------------
def masked_mse(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, coef: float,
               n_ref: int, divisor: float = 1.0, use_unmasked: bool = True) -> torch.Tensor:
    """mean((pred/divisor-target/divisor)^2) over mask==1 rows, plus ``coef`` times the same
    over mask==0 rows when mask.sum() != n_ref.  ``divisor`` = far for the depth term."""
    m1 = mask.reshape(-1) == 1
    m0 = mask.reshape(-1) == 0
    loss = torch.mean((pred[m1] / divisor - target[m1] / divisor) ** 2)
    if use_unmasked and float(mask.sum()) != n_ref:
        loss = loss + coef * torch.mean((pred[m0] / divisor - target[m0] / divisor) ** 2)
    return loss


def soft_mse(x: torch.Tensor, y: torch.Tensor, param, kind: int) -> torch.Tensor:
    """The soft-weighted losses, differentiable like the reference's lambdas (only (x - y) of the denominator is detached):
    kind 0  img2mse_softmask / img2mse_depth_softmask(x, y, temp)   NP/run_nerf_view.py:50,55
    kind 1  img2mse_softLpmask(x, y, coef)                          NP/run_nerf_view.py:58."""
    d = x - y
    if kind == 0:
        return torch.sum(torch.exp(d ** 2 / param) * d ** 2) / torch.sum(torch.exp(d.detach() ** 2 / param))
    w = d.abs() ** param + 1
    return torch.sum(w * d ** 2) / torch.sum(w).detach()


def mse_to_psnr(x: torch.Tensor) -> torch.Tensor:
    """NP/run_nerf_helpers.py:10."""
    return -10.0 * torch.log(x) / math.log(10.0)


# --------------------------------------------------------------------------------------
# summation order of torch.sum on the CPU (what NP/run_nerf_helpers.py:209 executes on the oracle device)
# --------------------------------------------------------------------------------------
def aten_cpu_sum_f32(x: np.ndarray, vec: int = 8) -> np.float32:
    """float32 sum of a contiguous row in the order of ATen's CPU kernel (SumKernel.cpp:
    vectorized_inner_sum -> row_sum -> multi_row_sum): ``vec``-wide lane partials with four interleaved
    accumulators and a 16-step cascade, then the scalar tail, then the lane partials in order.  K5 restates
    this order on the GPU (csrc/sampling.cu: aten_cpu_sum) so that cdf low bits -- and with them the
    searchsorted indices at near-ties -- equal the reference's.  Pure-Python loops: small cases only."""
    f32 = np.float32
    x = np.asarray(x, dtype=f32)

    def row_sum(get, size):
        size_ilp = size // 4
        acc = [[f32(0)] * 4 for _ in range(4)]
        i = 0
        while i + 16 <= size_ilp:
            for _ in range(16):
                for k in range(4):
                    acc[0][k] = f32(acc[0][k] + get(4 * i + k))
                i += 1
            for j in range(1, 4):
                for k in range(4):
                    acc[j][k] = f32(acc[j][k] + acc[j - 1][k])
                    acc[j - 1][k] = f32(0)
                if (i & (15 << (4 * j))) != 0:
                    break
        while i < size_ilp:
            for k in range(4):
                acc[0][k] = f32(acc[0][k] + get(4 * i + k))
            i += 1
        for j in range(1, 4):
            for k in range(4):
                acc[0][k] = f32(acc[0][k] + acc[j][k])
        p0 = acc[0][0]
        for r in range(size_ilp * 4, size):
            p0 = f32(p0 + get(r))
        for k in range(1, 4):
            p0 = f32(p0 + acc[0][k])
        return p0

    n = x.shape[0]
    if n < vec:
        return row_sum(lambda i: x[i], n)
    vs = n // vec
    parts = [row_sum(lambda v, j=j: x[v * vec + j], vs) for j in range(vec)]
    fin = f32(0)
    for k in range(vs * vec, n):
        fin = f32(fin + x[k])
    for j in range(vec):
        fin = f32(fin + parts[j])
    return fin
