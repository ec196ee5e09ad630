#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which is absent on the GPU
box):   python oracle/make_golden.py

The reference has no tests and no golden vectors of its own, so
its functions are imported read-only from ``/root/reference/nerf-pytorch-master`` and
executed on seeded inputs on the CPU; inputs and outputs are stored as small ``.npz``
files.  Modules the reference imports only for I/O (imageio, configargparse, lpips ...)
are absent from this image and replaced by empty stand-ins before the import; the two
helpers that hard-code ``.cuda()`` (run_nerf_view.py:596,622) are executed with
``Tensor.cuda`` patched to the identity.  No reference source is copied.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NP_DIR = "/root/reference/nerf-pytorch-master"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import nerf_oracle as O  # noqa: E402  (weights + harness inputs only)

ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
SMALL = dict(D=4, W=64, input_ch=63, input_ch_views=0, output_ch=5, skips=(2,), use_viewdirs=False)


def import_reference():
    for name in ["imageio", "ipdb", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "configargparse",
                 "tensorboardX", "pytorch_msssim", "lpips"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].__path__ = []          # let "import matplotlib.cm" resolve
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]

    class _LPIPS:
        def __init__(self, *a, **k):
            pass

        def to(self, *a, **k):
            return self

    sys.modules["lpips"].LPIPS = _LPIPS
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["pytorch_msssim"].ssim = lambda *a, **k: None
    sys.modules["pytorch_msssim"].ms_ssim = lambda *a, **k: None
    sys.modules["pytorch_msssim"].SSIM = object
    sys.modules["pytorch_msssim"].MS_SSIM = object
    sys.path.insert(0, NP_DIR)
    torch.cuda.current_device = lambda: 0
    import run_nerf_helpers as H
    import run_nerf as R
    try:
        import run_nerf_view as V
    except Exception as exc:  # pragma: no cover - reported, view goldens then skipped
        print("run_nerf_view import failed:", repr(exc))
        V = None
    return H, R, V


def load_into(module, params):
    sd = module.state_dict()
    for k in sd:
        sd[k] = params[k].clone().to(sd[k].dtype)
    module.load_state_dict(sd)
    return module


def npz(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print("wrote", name, {k: tuple(v.shape) for k, v in conv.items()})


def synthetic_rays(n, seed, dtype=torch.float32):
    """Workload A of SURVEY.md section 8d."""
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, 4.0]) + 0.1 * torch.randn(n, 3, generator=g)
    d = torch.tensor([0.0, 0.0, -1.0]) + 0.2 * torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return o.to(dtype), d.to(dtype)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    H, R, V = import_reference()

    # ---- Embedder / get_embedder --------------------------------------------------
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(64, 3, generator=g) * 12.0 - 6.0)
    e10, d10 = H.get_embedder(10, 0)
    e4, d4 = H.get_embedder(4, 0)
    npz("embed", x=x, e10=e10(x), e4=e4(x), dims=np.array([d10, d4]))

    # ---- NeRF module ------------------------------------------------------------------
    for tag, arch, seed in (("mlp_viewdirs", ARCH, 11), ("mlp_small_noview", SMALL, 12)):
        p = O.make_params(seed, **arch)
        net = load_into(H.NeRF(D=arch["D"], W=arch["W"], input_ch=arch["input_ch"],
                               input_ch_views=arch["input_ch_views"], output_ch=arch["output_ch"],
                               skips=list(arch["skips"]), use_viewdirs=arch["use_viewdirs"]), p)
        g = torch.Generator().manual_seed(seed)
        xin = torch.randn(96, arch["input_ch"] + arch["input_ch_views"], generator=g)
        with torch.no_grad():
            y32 = net(xin)
            y64 = net.double()(xin.double())
        npz(tag, x=xin, y=y32, y64=y64, seed=np.array(seed),
            keys=np.array(list(net.state_dict().keys())))

    # ---- raw2outputs ------------------------------------------------------------------
    g = torch.Generator().manual_seed(2)
    n, s = 40, 64
    raw = torch.randn(n, s, 4, generator=g) * 2.0
    z = torch.sort(2.0 + 4.0 * torch.rand(n, s, generator=g), dim=-1).values
    _, d = synthetic_rays(n, 3)
    d = d * 1.3
    for wb in (False, True):
        rgb, disp, acc, w, depth = (V or R).raw2outputs(raw, z, d, 0.0, wb) if V else (*R.raw2outputs(raw, z, d, 0.0, wb),)
        npz(f"raw2outputs_wb{int(wb)}", raw=raw, z=z, rays_d=d, rgb=rgb, disp=disp, acc=acc, weights=w, depth=depth)
    # pytest hook: uniform noise from numpy seed 0 (NP/run_nerf.py:290-294)
    rgb, disp, acc, w, depth = R.raw2outputs(raw, z, d, 1.0, True, pytest=True)
    np.random.seed(0)
    noise = torch.Tensor(np.random.rand(n, s) * 1.0)
    npz("raw2outputs_noise", raw=raw, z=z, rays_d=d, noise=noise, rgb=rgb, disp=disp, acc=acc, weights=w, depth=depth)

    # ---- sample_pdf (capture the internal searchsorted result) --------------------------
    g = torch.Generator().manual_seed(4)
    n = 48
    bins = torch.sort(2.0 + 4.0 * torch.rand(n, 63, generator=g), dim=-1).values
    wts = torch.rand(n, 62, generator=g) ** 4
    wts[:6] = 0.0                                   # empty rays: uniform pdf from the +1e-5
    wts[6:12, 10:] = 0.0
    captured = {}
    real_ss = torch.searchsorted

    def spy(cdf, u, **kw):
        out = real_ss(cdf, u, **kw)
        captured["cdf"], captured["u"], captured["inds"] = cdf.clone(), u.clone(), out.clone()
        return out

    H.torch.searchsorted = spy
    try:
        s_det = H.sample_pdf(bins, wts, 128, det=True)
        det_dbg = dict(captured)
        s_rnd = H.sample_pdf(bins, wts, 128, det=False, pytest=True)
        rnd_dbg = dict(captured)
    finally:
        H.torch.searchsorted = real_ss
    npz("sample_pdf", bins=bins, weights=wts, samples_det=s_det, cdf_det=det_dbg["cdf"], u_det=det_dbg["u"],
        inds_det=det_dbg["inds"], samples_rnd=s_rnd, cdf_rnd=rnd_dbg["cdf"], u_rnd=rnd_dbg["u"],
        inds_rnd=rnd_dbg["inds"])

    # ---- render_rays / render end to end ------------------------------------------------
    n, S, NI = 40, 64, 128
    o, d = synthetic_rays(n, 5)
    pc = O.make_params(21, sigma_bias=0.15, **ARCH)
    pf = O.make_params(22, sigma_bias=0.15, **ARCH)
    mod = V or R
    coarse = load_into(H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True), pc)
    fine = load_into(H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True), pf)
    query = lambda inputs, viewdirs, fn: mod.run_network(inputs, viewdirs, fn, embed_fn=e10, embeddirs_fn=e4,
                                                         netchunk=1024 * 64)
    kw = dict(network_query_fn=query, perturb=0.0, N_importance=NI, network_fine=fine, N_samples=S,
              network_fn=coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False)
    with torch.no_grad():
        ret = mod.render(8, 8, np.eye(3), chunk=16, rays=torch.stack([o, d]), near=2.0, far=6.0, retraw=True, **kw)
    names = ["rgb_map", "disp_map", "acc_map"] + (["depth_map"] if V else [])
    out = {k: v for k, v in zip(names, ret[:-1])}
    out.update(ret[-1])
    npz("render_det", rays_o=o, rays_d=d, near=np.array(2.0), far=np.array(6.0), seeds=np.array([21, 22]),
        sigma_bias=np.array(0.15), **out)

    # stochastic branch through the pytest hook (perturb=1, raw_noise_std=1, lindisp)
    kw2 = dict(kw, perturb=1.0, raw_noise_std=1.0, lindisp=True, white_bkgd=False, pytest=True)
    with torch.no_grad():
        ret = mod.render(8, 8, np.eye(3), chunk=1024, rays=torch.stack([o, d]), near=2.0, far=6.0, retraw=True, **kw2)
    out = {k: v for k, v in zip(names, ret[:-1])}
    out.update(ret[-1])
    np.random.seed(0); t_rand = np.random.rand(n, S).astype(np.float32)
    np.random.seed(0); noise_c = (np.random.rand(n, S) * 1.0).astype(np.float32)
    np.random.seed(0); u = np.random.rand(n, NI).astype(np.float32)
    np.random.seed(0); noise_f = (np.random.rand(n, S + NI) * 1.0).astype(np.float32)
    npz("render_pytest", rays_o=o, rays_d=d, near=np.array(2.0), far=np.array(6.0), seeds=np.array([21, 22]),
        sigma_bias=np.array(0.15), t_rand=t_rand, noise_c=noise_c, u=u, noise_f=noise_f, **out)

    # ---- get_rays / ndc_rays / render() ray packing -------------------------------------
    Hh, Ww, focal = 6, 8, 7.5
    K = np.array([[focal, 0, 0.5 * Ww], [0, focal, 0.5 * Hh], [0, 0, 1]])
    c2w = torch.tensor([[0.9, -0.1, 0.42, 0.3], [0.2, 0.95, -0.2, -0.1], [-0.38, 0.27, 0.88, 3.5]])
    ro, rd = H.get_rays(Hh, Ww, K, c2w)
    no, nd = H.ndc_rays(Hh, Ww, focal, 1.0, ro, rd)
    npz("rays", K=K, c2w=c2w, rays_o=ro.contiguous(), rays_d=rd, ndc_o=no, ndc_d=nd, hw=np.array([Hh, Ww]),
        focal=np.array(focal))

    # ---- cross-view geometry: get_ref_rays / get_test_label + hard-mask rule ---------------
    if V is not None:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.LongTensor = torch.LongTensor
        Hh, Ww, focal = 24, 32, 30.0
        Kt = torch.tensor([[focal, 0, 0.5 * Ww], [0, focal, 0.5 * Hh], [0, 0, 1.0]])

        def pose(tx, ry):
            c, s = np.cos(ry), np.sin(ry)
            m = np.eye(4, dtype=np.float32)
            m[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
            m[:3, 3] = [tx, 0.05, 4.0]
            return torch.from_numpy(m)

        c2w_t, c2w_r = pose(0.0, 0.0), pose(0.35, 0.08)
        w2c_r = torch.inverse(c2w_r)
        g = torch.Generator().manual_seed(7)
        img_r = torch.rand(1, 3, Hh, Ww, generator=g)
        yy, xx = torch.meshgrid(torch.arange(Hh, dtype=torch.float32), torch.arange(Ww, dtype=torch.float32), indexing="ij")
        depth_t = 3.6 + 0.4 * torch.sin(xx / 5.0) * torch.cos(yy / 4.0)
        depth_r = 3.6 + 0.4 * torch.sin((xx + 2.6) / 5.0) * torch.cos(yy / 4.0) + 0.02 * torch.randn(Hh, Ww, generator=g)
        ro, rd = H.get_rays(Hh, Ww, Kt.numpy(), c2w_t[:3, :4])
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        pw = ro + depth_t.reshape(-1, 1) * rd
        rgb_ref, dep_ref, cam, r_o, r_d, inb = V.get_ref_rays(w2c_r[None], c2w_r[None], Kt[None], pw[None, :, None, :],
                                                               img_r, depth_r[None])
        ty, tx, tmask, tz = V.get_test_label(w2c_r[None], c2w_r[None], Kt[None], pw[None, :, None, :], img_r)
        # the hard-mask rule of train() (run_nerf_view.py:1014-1039), driven through the
        # reference's own get_ref_rays, chunk = 200 pixels to exercise the doubling
        chunk, thr0 = 200, 0.01
        mask_ref = torch.zeros(Hh * Ww, dtype=torch.bool)
        for c0 in range(0, Hh * Ww, chunk):
            sl = slice(c0, c0 + chunk)
            _, dr, pc_, _, _, mb = V.get_ref_rays(w2c_r[None], c2w_r[None], Kt[None], pw[None, sl, None, :], img_r, depth_r[None])
            if mb.sum() != 0:
                m = torch.ones(dr.shape[2], 1) < 0
                thr = thr0
                while m.sum() == 0:
                    diff = pc_[mb][..., -1].unsqueeze(-1) - dr.squeeze(0).squeeze(0)[:, None]
                    m = abs(diff) < thr
                    thr = 2 * thr
                mb2 = mb.clone()
                mb2[mb] = m.squeeze()
                mask_ref[sl] = mb2.squeeze()
        npz("crossview", K=Kt, c2w_tgt=c2w_t, c2w_ref=c2w_r, w2c_ref=w2c_r, img_ref=img_r[0], depth_tgt=depth_t,
            depth_ref=depth_r, rays_o=ro, rays_d=rd, pts_w=pw, rgb_ref=rgb_ref[0].t(), dep_ref=dep_ref.reshape(-1),
            cam=cam[0], ref_rays_o=r_o, ref_rays_d=r_d, inb=inb[0], label_y=ty[0], label_x=tx[0], label_mask=tmask[0],
            label_z=tz[0], hard_mask=mask_ref, chunk=np.array(chunk), thr0=np.array(thr0))

        # ---- masked losses (run_nerf_view.py:1645-1648, 1737; cal_correspondance 1550-1551) ------
        g = torch.Generator().manual_seed(9)
        nr = 300
        rgb = torch.rand(nr, 3, generator=g); tgt = torch.rand(nr, 3, generator=g)
        dp = 2 + 4 * torch.rand(nr, generator=g); dq = 2 + 4 * torch.rand(nr, generator=g)
        m = (torch.rand(nr, 1, generator=g) > 0.4).float()
        far, coef = 6.0, 0.2
        li = V.img2mse(rgb[m.squeeze() == 1], tgt[m.squeeze() == 1])
        if m.squeeze().sum() != nr:
            li = li + coef * V.img2mse(rgb[m.squeeze() == 0], tgt[m.squeeze() == 0])
        ld1 = V.img2mse(dp[m.squeeze() == 1] / far, dq[m.squeeze() == 1] / far)
        ld2 = ld1 + coef * V.img2mse(dp[m.squeeze() == 0] / far, dq[m.squeeze() == 0] / far)
        npz("masked_loss", rgb=rgb, tgt=tgt, depth=dp, depth_prior=dq, mask=m, far=np.array(far), coef=np.array(coef),
            img_loss=li, depth_loss_masked_only=ld1, depth_loss_both=ld2, psnr=V.mse2psnr(li))


if __name__ == "__main__":
    main()
