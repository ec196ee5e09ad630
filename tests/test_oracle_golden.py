"""Pins the CPU oracle against vectors produced by the unmodified reference
(oracle/make_golden.py).  float32 on CPU reproduces the reference arithmetic, so most
checks are exact or at round-off."""
import numpy as np
import torch

from conftest import load_golden, t
from oracle import nerf_oracle as O

ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
SMALL = dict(D=4, W=64, input_ch=63, input_ch_views=0, output_ch=5, skips=(2,), use_viewdirs=False)


def close(a, b, rtol=1e-6, atol=1e-7):
    a = a.numpy() if isinstance(a, torch.Tensor) else a
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def test_posenc_matches_embedder():
    g = load_golden("embed")
    x = t(g["x"])
    assert tuple(g["dims"]) == (63, 27)
    assert torch.equal(O.posenc(x, 10), t(g["e10"]))
    assert torch.equal(O.posenc(x, 4), t(g["e4"]))


def test_param_keys_and_mlp():
    for tag, arch in (("mlp_viewdirs", ARCH), ("mlp_small_noview", SMALL)):
        g = load_golden(tag)
        keys = [k for k, _ in O.nerf_param_shapes(**arch)]
        assert keys == [str(k) for k in g["keys"]]
        p = O.make_params(int(g["seed"]), **arch)
        y = O.mlp_forward(p, t(g["x"]), **arch)
        close(y, g["y"], rtol=2e-5, atol=2e-6)
        p64 = O.make_params(int(g["seed"]), dtype=torch.float64, **arch)
        y64 = O.mlp_forward(p64, t(g["x"]).double(), **arch)
        # reference .double() keeps fp32-rounded weights; ours are drawn in fp64 -> loose
        close(y64, g["y64"], rtol=1e-4, atol=1e-5)


def test_composite_matches_raw2outputs():
    for tag, wb in (("raw2outputs_wb0", False), ("raw2outputs_wb1", True), ("raw2outputs_noise", True)):
        g = load_golden(tag)
        noise = t(g["noise"]) if "noise" in g else None
        c = O.composite(t(g["raw"]), t(g["z"]), t(g["rays_d"]), noise, wb)
        for k in ("rgb", "disp", "acc", "weights", "depth"):
            assert torch.equal(c[k], t(g[k])), (tag, k)


def test_sample_pdf_indices_bit_exact():
    g = load_golden("sample_pdf")
    for mode in ("det", "rnd"):
        s, dbg = O.sample_pdf(t(g["bins"]), t(g["weights"]), t(g["u_" + mode]), return_debug=True)
        assert torch.equal(dbg["cdf"], t(g["cdf_" + mode]))
        inds = t(g["inds_" + mode])
        assert torch.equal(dbg["below"], torch.clamp(inds - 1, min=0))
        assert torch.equal(dbg["above"], torch.clamp(inds, max=62))
        assert torch.equal(s, t(g["samples_" + mode]))


def _render(g, stochastic):
    o, d = t(g["rays_o"]), t(g["rays_d"])
    rays = O.pack_rays(o, d, float(g["near"]), float(g["far"]), True)
    pc = O.make_params(int(g["seeds"][0]), sigma_bias=float(g["sigma_bias"]), **ARCH)
    pf = O.make_params(int(g["seeds"][1]), sigma_bias=float(g["sigma_bias"]), **ARCH)
    kw = dict(n_samples=64, n_importance=128, retraw=True)
    if stochastic:
        kw.update(lindisp=True, white_bkgd=False, t_rand=t(g["t_rand"]), u=t(g["u"]),
                  noise_coarse=t(g["noise_c"]), noise_fine=t(g["noise_f"]))
    else:
        kw.update(white_bkgd=True)
    return O.render_rays(rays, pc, pf, ARCH, **kw)


def test_render_rays_det_and_pytest_hook():
    for tag, stochastic in (("render_det", False), ("render_pytest", True)):
        g = load_golden(tag)
        out = _render(g, stochastic)
        for k in ("rgb_map", "acc_map", "depth_map", "rgb0", "acc0", "depth0", "z_std", "raw"):
            close(out[k], g[k], rtol=3e-5, atol=3e-6)
        close(out["disp_map"], g["disp_map"], rtol=1e-4, atol=1e-6)


def test_rays_and_ndc():
    g = load_golden("rays")
    H, W = (int(v) for v in g["hw"])
    ro, rd = O.pixel_rays(H, W, g["K"], t(g["c2w"]))
    close(ro, g["rays_o"], rtol=0, atol=0)
    close(rd, g["rays_d"], rtol=1e-6)
    no, nd = O.ndc_warp(H, W, float(g["focal"]), 1.0, t(g["rays_o"]), t(g["rays_d"]))
    close(no, g["ndc_o"]); close(nd, g["ndc_d"])


def test_crossview_projection_gather_and_hard_mask():
    g = load_golden("crossview")
    H, W = g["depth_ref"].shape
    px, py, inb, cam = O.project_points(t(g["pts_w"]), t(g["w2c_ref"]), t(g["K"]), H, W)
    assert np.array_equal(px.numpy(), g["label_x"]) and np.array_equal(py.numpy(), g["label_y"])
    assert np.array_equal(inb.numpy(), g["inb"]) and np.array_equal(inb.numpy(), g["label_mask"])
    close(cam, g["cam"], rtol=2e-6, atol=2e-6)
    close(cam[:, 2], g["label_z"], rtol=2e-6, atol=2e-6)
    rgb, dep = O.gather_reference(t(g["img_ref"]), t(g["depth_ref"]), px, py, inb)
    assert np.array_equal(rgb[inb].numpy(), g["rgb_ref"]) and np.array_equal(dep[inb].numpy(), g["dep_ref"])
    ro, rd = O.reference_view_rays(px[inb], py[inb], t(g["K"]), t(g["c2w_ref"]))
    close(ro, g["ref_rays_o"]); close(rd, g["ref_rays_d"], rtol=1e-6, atol=1e-7)
    hm = O.hard_mask_pair(t(g["rays_o"]), t(g["rays_d"]), t(g["depth_tgt"]).reshape(-1), t(g["w2c_ref"]),
                          t(g["K"]), t(g["depth_ref"]), thr0=float(g["thr0"]), chunk=int(g["chunk"]))
    assert np.array_equal(hm.numpy(), g["hard_mask"])
    assert 0 < hm.sum() < hm.numel()


def test_masked_losses():
    g = load_golden("masked_loss")
    rgb, tgt, m = t(g["rgb"]), t(g["tgt"]), t(g["mask"])
    far, coef = float(g["far"]), float(g["coef"])
    li = O.masked_mse(rgb, tgt, m, coef, n_ref=rgb.shape[0])
    close(li, g["img_loss"], rtol=1e-6)
    close(O.mse_to_psnr(li), g["psnr"].reshape(()), rtol=1e-6)
    d, q = t(g["depth"]), t(g["depth_prior"])
    close(O.masked_mse(d, q, m, coef, n_ref=d.shape[0], divisor=far, use_unmasked=False),
          g["depth_loss_masked_only"], rtol=1e-6)
    close(O.masked_mse(d, q, m, coef, n_ref=d.shape[0], divisor=far), g["depth_loss_both"], rtol=1e-6)


def test_sum_order_restatement_matches_torch_cpu():
    """K5's total is summed in ATen's CPU order; this pins the restatement of that order to torch.sum here."""
    rs = np.random.RandomState(0)
    for n in (1, 3, 7, 8, 9, 31, 61, 62, 63, 64, 127, 129, 255, 600, 1031):
        X = (rs.rand(16, n) ** 3).astype(np.float32)
        ref = torch.from_numpy(X).sum(-1, keepdim=True).numpy()[:, 0]
        mine = np.array([O.aten_cpu_sum_f32(X[r]) for r in range(16)], dtype=np.float32)
        assert np.array_equal(ref, mine), n
    g = load_golden("sample_pdf")
    w = (g["weights"] + np.float32(1e-5)).astype(np.float32)
    tot = np.array([O.aten_cpu_sum_f32(r) for r in w], dtype=np.float32)
    assert np.array_equal(tot, torch.from_numpy(w).sum(-1).numpy())


# ---- round-2 goldens (oracle/make_golden_r2.py): NDC render, coarse-only / no view directions, hard mask at chunk 5120 ----
NOVIEW = dict(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=4, skips=(4,), use_viewdirs=False)


def test_render_ndc_golden():
    g = load_golden("render_ndc")
    H, W = (int(v) for v in g["hw"])
    ro, rd = O.pixel_rays(H, W, g["K"], t(g["c2w"]))
    close(ro.reshape(-1, 3), g["rays_o"], rtol=0, atol=0)
    rays = O.pack_rays(ro.reshape(-1, 3), rd.reshape(-1, 3), 0.0, 1.0, True, ndc=(H, W, float(g["focal"])))
    pc = O.make_params(int(g["seeds"][0]), sigma_bias=float(g["sigma_bias"]), **ARCH)
    pf = O.make_params(int(g["seeds"][1]), sigma_bias=float(g["sigma_bias"]), **ARCH)
    out = O.render_rays(rays, pc, pf, ARCH, n_samples=64, n_importance=128, white_bkgd=False, retraw=True)
    for k in ("rgb_map", "acc_map", "depth_map", "rgb0", "acc0", "depth0", "z_std"):
        close(out[k], g[k].reshape(H * W, *g[k].shape[2:]), rtol=3e-5, atol=3e-6)
    close(out["raw"], g["raw"].reshape(H * W, 192, 4), rtol=3e-5, atol=3e-6)


def test_render_noview_coarse_only_golden():
    g = load_golden("render_noview")
    rays = O.pack_rays(t(g["rays_o"]), t(g["rays_d"]), float(g["near"]), float(g["far"]), False)
    assert rays.shape[1] == 8
    p = O.make_params(int(g["seed"]), sigma_bias=float(g["sigma_bias"]), **NOVIEW)
    out = O.render_rays(rays, p, None, NOVIEW, n_samples=32, n_importance=0, white_bkgd=True, retraw=True)
    for k in ("rgb_map", "acc_map", "raw"):
        close(out[k], g[k], rtol=3e-5, atol=3e-6)
    close(out["disp_map"], g["disp_map"], rtol=1e-4, atol=1e-6)
    assert "rgb0" not in out


def test_hard_mask_at_the_real_chunk_size_golden():
    g = load_golden("hardmask_big")
    n = g["mask"].shape[0]
    assert n == 16128 and int(g["chunk"]) == 5120 and n // 5120 == 3          # three full chunks + a ragged one
    total = torch.zeros(n, dtype=torch.bool)
    for r in range(2):
        m = O.hard_mask_pair(t(g["rays_o"]), t(g["rays_d"]), t(g["depth_tgt"]).reshape(-1), t(g["w2c_refs"][r]), t(g["K"]),
                             t(g["depth_refs"][r]), thr0=float(g["thr0"]), chunk=5120)
        assert np.array_equal(m.numpy(), g[f"mask_ref{r}"]), r
        total |= m
    assert np.array_equal(total.numpy(), g["mask"]) and 0 < int(total.sum()) < n
    # the chunk size is a semantic parameter: a different chunking gives a different mask on this scene
    other = O.hard_mask_pair(t(g["rays_o"]), t(g["rays_d"]), t(g["depth_tgt"]).reshape(-1), t(g["w2c_refs"][0]), t(g["K"]),
                             t(g["depth_refs"][0]), thr0=float(g["thr0"]), chunk=16128)
    assert not np.array_equal(other.numpy(), g["mask_ref0"])
