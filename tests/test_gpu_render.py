"""End-to-end parity of render / render_rays (the drop-in surface) against the reference's golden
vectors and the fp64 oracle, plus training-mode gradients against the oracle's autograd."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import nerf_oracle as O
from util import ARCH, assert_close, module_from_params, rel_err, workload_rays

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cn():
    import consistentnerf_b200 as m
    return m


def _kwargs(cn, coarse, fine, **over):
    embed_fn, _ = cn.get_embedder(10, 0)
    embeddirs_fn, _ = cn.get_embedder(4, 0)

    def network_query_fn(inputs, viewdirs, network_fn):      # create_nerf's closure, NP/run_nerf_view.py:323-326
        return cn.run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
                              netchunk=1024 * 64)
    kw = dict(network_query_fn=network_query_fn, perturb=0.0, N_importance=128, network_fine=fine, N_samples=64,
              network_fn=coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False,
              near=2.0, far=6.0)
    kw.update(over)
    return kw


def _nets(g):
    pc = O.make_params(int(g["seeds"][0]), sigma_bias=float(g["sigma_bias"]), **ARCH)
    pf = O.make_params(int(g["seeds"][1]), sigma_bias=float(g["sigma_bias"]), **ARCH)
    return pc, pf, module_from_params(pc, ARCH), module_from_params(pf, ARCH)


def test_render_deterministic_golden(cn):
    g = load_golden("render_det")
    pc, pf, coarse, fine = _nets(g)
    rays = (t(g["rays_o"], device=DEV), t(g["rays_d"], device=DEV))
    kw = _kwargs(cn, coarse, fine, near=float(g["near"]), far=float(g["far"]))
    with torch.no_grad():
        rgb, disp, acc, depth, extras = cn.render(40, 40, None, chunk=16, rays=rays, retraw=True, **kw)
    got = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, **extras)
    for k in ("rgb_map", "acc_map", "depth_map", "rgb0", "acc0", "depth0", "z_std", "raw"):
        assert_close(got[k], g[k], 1e-4, 1e-5, k)
    assert_close(got["disp_map"], g["disp_map"], 2e-4, 1e-6, "disp_map")
    # chunk invariance (batchify_rays contract, NP/run_nerf.py:79-80)
    with torch.no_grad():
        again = cn.render(40, 40, None, chunk=1024, rays=rays, retraw=True, **kw)
    assert torch.equal(again[0], rgb) and torch.equal(again[3], depth)


def test_render_stochastic_pytest_hook_golden(cn):
    """perturb=1, raw_noise_std=1, lindisp, black background, under the reference's fixed-RNG hook."""
    g = load_golden("render_pytest")
    pc, pf, coarse, fine = _nets(g)
    rays = (t(g["rays_o"], device=DEV), t(g["rays_d"], device=DEV))
    kw = _kwargs(cn, coarse, fine, near=float(g["near"]), far=float(g["far"]), perturb=1.0, raw_noise_std=1.0,
                 lindisp=True, white_bkgd=False, pytest=True)
    with torch.no_grad():
        rgb, disp, acc, depth, extras = cn.render(40, 40, None, chunk=1024, rays=rays, retraw=True, **kw)
    got = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, **extras)
    for k in ("rgb_map", "acc_map", "depth_map", "rgb0", "acc0", "depth0", "z_std", "raw"):
        assert_close(got[k], g[k], 1e-4, 1e-5, k)


def test_vanilla_flavour_has_no_depth(cn):
    g = load_golden("render_det")
    pc, pf, coarse, fine = _nets(g)
    api = cn.make_api(with_depth=False)
    rays = (t(g["rays_o"], device=DEV), t(g["rays_d"], device=DEV))
    with torch.no_grad():
        out = api.render(40, 40, None, chunk=1024, rays=rays, **_kwargs(cn, coarse, fine))
    assert len(out) == 4 and "depth0" not in out[3] and "rgb0" in out[3]
    assert_close(out[0], g["rgb_map"], 1e-4, 1e-5)


def test_render_workload_a_vs_fp64_oracle(cn):
    """512 rays of the headline workload, judged against the fp64 restatement:
    the candidate must be within 1e-4 (scale-relative) wherever the fp32 oracle itself is."""
    n = 512
    o, d = workload_rays(n)
    pc = O.make_params(0, sigma_bias=0.5, **ARCH)
    pf = O.make_params(1, sigma_bias=0.5, **ARCH)
    coarse, fine = module_from_params(pc, ARCH), module_from_params(pf, ARCH)
    with torch.no_grad():
        rgb, disp, acc, depth, ex = cn.render(1, n, None, chunk=4096, rays=(o.to(DEV), d.to(DEV)),
                                              **_kwargs(cn, coarse, fine))
    rays64 = O.pack_rays(o.double(), d.double(), 2.0, 6.0, True)
    ref = O.render_rays(rays64, {k: v.double() for k, v in pc.items()}, {k: v.double() for k, v in pf.items()}, ARCH,
                        n_samples=64, n_importance=128, white_bkgd=True)
    ref32 = O.render_rays(rays64.float(), pc, pf, ARCH, n_samples=64, n_importance=128, white_bkgd=True)
    got = dict(rgb_map=rgb, acc_map=acc, depth_map=depth, rgb0=ex["rgb0"], depth0=ex["depth0"], acc0=ex["acc0"])
    for k, v in got.items():
        e, e32 = rel_err(v, ref[k]), rel_err(ref32[k], ref[k])
        print(f"{k:10s} candidate {e:.2e}   fp32 oracle {e32:.2e}")
        assert e < max(1e-4, 2 * e32), k


def test_training_gradients_match_oracle_autograd(cn):
    """loss = img2mse(rgb, tgt) + img2mse(rgb0, tgt) (NP/run_nerf.py:770-777), perturb + noise supplied
    through the pytest hook so that oracle and kernels see identical random numbers."""
    n = 96
    o, d = workload_rays(n, seed=3)
    pc = O.make_params(2, sigma_bias=0.5, **ARCH)
    pf = O.make_params(3, sigma_bias=0.5, **ARCH)
    coarse, fine = module_from_params(pc, ARCH), module_from_params(pf, ARCH)
    tgt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    kw = _kwargs(cn, coarse, fine, perturb=1.0, raw_noise_std=1.0, pytest=True)
    rgb, disp, acc, depth, ex = cn.render(1, n, None, chunk=4096, rays=(o.to(DEV), d.to(DEV)), retraw=True, **kw)
    loss = cn.img2mse(rgb, tgt.to(DEV)) + cn.img2mse(ex["rgb0"], tgt.to(DEV))
    loss.backward()

    np.random.seed(0); t_rand = torch.tensor(np.random.rand(n, 64))
    np.random.seed(0); noise_c = torch.tensor(np.random.rand(n, 64))
    np.random.seed(0); u = torch.tensor(np.random.rand(n, 128))
    np.random.seed(0); noise_f = torch.tensor(np.random.rand(n, 192))
    # the kernels consume fp32 random numbers: round the oracle's copies the same way
    rnd = dict(t_rand=t_rand.float(), u=u.float(), noise_coarse=noise_c.float(), noise_fine=noise_f.float())

    def oracle_grads(dt):
        c = {k: v.to(dt).requires_grad_(True) for k, v in pc.items()}
        f = {k: v.to(dt).requires_grad_(True) for k, v in pf.items()}
        rays = O.pack_rays(o.to(dt), d.to(dt), 2.0, 6.0, True)
        ref = O.render_rays(rays, c, f, ARCH, n_samples=64, n_importance=128, white_bkgd=True,
                            **{k: v.to(dt) for k, v in rnd.items()})
        ref_loss = ((ref["rgb_map"] - tgt.to(dt)) ** 2).mean() + ((ref["rgb0"] - tgt.to(dt)) ** 2).mean()
        ref_loss.backward()
        return float(ref_loss.detach()), c, f

    loss64, c64, f64 = oracle_grads(torch.float64)
    _, c32, f32 = oracle_grads(torch.float32)
    assert abs(float(loss.detach()) - loss64) < 1e-4 * abs(loss64)
    # Adjudication: the fp32 reference path itself is ~1e-3 away from fp64 on the early
    # layers (fp32 sample positions amplified by the 2^9 octave), so the candidate is held to twice the fp32
    # oracle's own distance from fp64, with a floor of 5e-4 of the largest entry: the density head's gradient is
    # a heavily cancelling sum (max |g| ~1e-6 out of terms ~1e-4) where summation order alone moves 2e-4.
    for net, p64, p32 in ((coarse, c64, c32), (fine, f64, f32)):
        for name, prm in net.named_parameters():
            if p64[name].grad is None:
                assert prm.grad is None or float(prm.grad.abs().max()) == 0.0
                continue
            cand, floor = rel_err(prm.grad, p64[name].grad), rel_err(p32[name].grad, p64[name].grad)
            assert cand < max(5e-4, 2.0 * floor), (name, cand, floor)


def test_gradients_accumulate_into_existing_grad_buffers(cn):
    """Parameters that already own a .grad (FlatGrads, or zero_grad(set_to_none=False)): the reduction kernels add into it and
    autograd gets no gradient for them; the result must equal what AccumulateGrad produces from returned tensors."""
    from consistentnerf_b200.distributed import FlatGrads
    n = 64
    o, d = workload_rays(n, seed=4)
    pc, pf = O.make_params(2, sigma_bias=0.5, **ARCH), O.make_params(3, sigma_bias=0.5, **ARCH)
    tgt = torch.rand(n, 3, generator=torch.Generator().manual_seed(6)).to(DEV)

    def run(in_place, passes):
        coarse, fine = module_from_params(pc, ARCH), module_from_params(pf, ARCH)
        hot = [p for net in (coarse, fine) for nm, p in net.named_parameters() if nm in net.spec.param_names()]
        if in_place:
            flat = FlatGrads(hot)
            flat.flat.fill_(2.0 ** -27)           # pre-existing content (of the gradients' own magnitude) must be kept and added to
        kw = _kwargs(cn, coarse, fine)
        for _ in range(passes):
            rgb, disp, acc, depth, ex = cn.render(1, n, None, chunk=4096, rays=(o.to(DEV), d.to(DEV)), retraw=True, **kw)
            (cn.img2mse(rgb, tgt) + cn.img2mse(ex["rgb0"], tgt)).backward()
        return [p.grad.clone() for p in hot]

    ref, got = run(False, 2), run(True, 2)
    for a, b in zip(ref, got):
        assert rel_err(b - 2.0 ** -27, a) < 1e-5


def test_whole_image_render_from_pose(cn):
    """render(c2w=...) path of render_path (NP/run_nerf_view.py:270): rays generated on the device."""
    H = W = 20
    K = np.array([[25.0, 0, 10.0], [0, 25.0, 10.0], [0, 0, 1]], dtype=np.float32)
    c2w = torch.tensor([[1.0, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 4.0]])
    pc = O.make_params(4, sigma_bias=0.5, **ARCH)
    pf = O.make_params(5, sigma_bias=0.5, **ARCH)
    coarse, fine = module_from_params(pc, ARCH), module_from_params(pf, ARCH)
    with torch.no_grad():
        rgb, disp, acc, depth, _ = cn.render(H, W, K, chunk=128, c2w=c2w.to(DEV), **_kwargs(cn, coarse, fine))
    assert rgb.shape == (H, W, 3) and depth.shape == (H, W)
    ro, rd = O.pixel_rays(H, W, K, c2w)
    rays = O.pack_rays(ro.reshape(-1, 3), rd.reshape(-1, 3), 2.0, 6.0, True)
    ref = O.render_rays(rays, pc, pf, ARCH, n_samples=64, n_importance=128, white_bkgd=True)
    assert_close(rgb.reshape(-1, 3), ref["rgb_map"], 1e-4, 1e-5)
    assert_close(depth.reshape(-1), ref["depth_map"], 1e-4, 1e-5)


def test_ray_bank_batches_match_reference_sampling(cn):
    """The device batch sampler yields exactly what train() builds per step (NP/run_nerf_view.py:1443-1517): rays through
    the chosen pixels of the chosen view and the targets / prior depths / masks gathered at those pixels."""
    V, H, W = 3, 12, 20
    gen = torch.Generator().manual_seed(7)
    images = torch.rand(V, H, W, 3, generator=gen)
    depths = 2 + torch.rand(V, H, W, generator=gen)
    masks = (torch.rand(V, H, W, generator=gen) > 0.5).float()
    K = np.array([[30.0, 0, W / 2], [0, 30.0, H / 2], [0, 0, 1]], dtype=np.float32)
    poses = torch.eye(4)[None].repeat(V, 1, 1)
    poses[:, :3, 3] = torch.randn(V, 3, generator=gen)
    poses[1, :3, :3] = torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    bank = cn.RayBank(images.numpy(), poses.numpy(), K, near=2.0, far=6.0, use_viewdirs=True, depths=depths.numpy(),
                      masks=masks.numpy(), seed=3)
    for kw in (dict(n_rand=64), dict(n_rand=50, patches=2, patch_size=4), dict(n_rand=30, precrop_frac=0.5, view=1)):
        b = bank.sample(**kw)
        v, pix = b["view"], b["pix"].cpu().long()
        n = kw["n_rand"] + kw.get("patches", 0) * kw.get("patch_size", 16) ** 2
        assert pix.shape[0] == n and b["rays"].shape == (n, 11)
        rand_part = pix[-kw["n_rand"]:]
        assert rand_part.unique().numel() == kw["n_rand"]                     # replace=False
        if "precrop_frac" in kw:
            yy, xx = rand_part // W, rand_part % W
            assert yy.min() >= H // 2 - 3 and yy.max() <= H // 2 + 2 and xx.min() >= W // 2 - 5 and xx.max() <= W // 2 + 4
        ro, rd = O.pixel_rays(H, W, K, poses[v, :3, :4])
        ref = O.pack_rays(ro.reshape(-1, 3)[pix], rd.reshape(-1, 3)[pix], 2.0, 6.0, True)
        assert_close(b["rays"], ref, 1e-6, 1e-7)
        assert torch.equal(b["target"].cpu(), images[v].reshape(-1, 3)[pix])
        assert torch.equal(b["depth"].cpu(), depths[v].reshape(-1)[pix])
        assert torch.equal(b["mask"].cpu(), masks[v].reshape(-1)[pix])
        assert torch.equal(b["batch_rays"][1], b["rays"][:, 3:6])


def test_render_path_matches_per_image_render(cn):
    H = W = 16
    K = np.array([[20.0, 0, 8.0], [0, 20.0, 8.0], [0, 0, 1]], dtype=np.float32)
    poses = []
    for k in range(3):
        c2w = torch.eye(4)
        c2w[:3, 3] = torch.tensor([0.1 * k, -0.1, 4.0])
        poses.append(c2w)
    pc = O.make_params(4, sigma_bias=0.5, **ARCH)
    pf = O.make_params(5, sigma_bias=0.5, **ARCH)
    coarse, fine = module_from_params(pc, ARCH), module_from_params(pf, ARCH)
    kw = _kwargs(cn, coarse, fine)
    rgbs, disps, accs = cn.render_path(torch.stack(poses).to(DEV), (H, W, 20.0), K, 4096, kw)
    assert rgbs.shape == (3, H, W, 3) and disps.shape == (3, H, W) and accs.shape == (3, H, W)
    with torch.no_grad():
        for k in range(3):
            rgb, disp, acc, depth, _ = cn.render(H, W, K, chunk=4096, c2w=poses[k][:3, :4].to(DEV), **kw)
            assert np.array_equal(rgbs[k], rgb.cpu().numpy()) and np.array_equal(accs[k], acc.cpu().numpy())
            np.testing.assert_array_equal(disps[k], disp.cpu().numpy())


def test_step_log_and_loss_scalars_without_per_step_sync(cn):
    """The per-step logging of train() (NP/run_nerf_view.py:1908-1937): the logged quantities come out of the loss kernel's
    own statistics and reach the host once per flush; values must equal the reference formulas."""
    g = torch.Generator().manual_seed(21)
    log = cn.StepLog(["loss", "mse", "psnr", "masked_psnr", "n_masked"], capacity=8)
    expect = []
    for step in range(5):
        rgb, tgt = torch.rand(300, 3, generator=g), torch.rand(300, 3, generator=g)
        mask = (torch.rand(300, 1, generator=g) > 0.4).float()
        loss, stats = cn.masked_img_loss(rgb.to(DEV), tgt.to(DEV), mask.to(DEV), 0.2, return_stats=True)
        log.record(step, loss=loss, not_logged=1.0, **cn.loss_scalars(stats))
        mse = ((rgb - tgt) ** 2).mean()
        m1 = mask[:, 0] == 1
        ref_loss = O.masked_mse(rgb, tgt, mask, 0.2, 300)
        expect.append(dict(loss=float(ref_loss), mse=float(mse), psnr=float(-10.0 * torch.log(mse) / np.log(10.0)),
                           n_masked=float(m1.sum())))
    rows = log.flush()
    assert [s for s, _ in rows] == list(range(5)) and log.flush() == []
    for (_, got), ref in zip(rows, expect):
        for k, v in ref.items():
            assert abs(got[k] - v) <= 2e-5 * max(1.0, abs(v)), (k, got[k], v)
        assert np.isfinite(got["masked_psnr"])


@pytest.mark.parametrize("n", [0, 1, 129])
def test_render_edge_sizes(cn, n):
    """Empty, single-ray and just-over-one-tile batches through the whole path (forward and backward): shapes follow the
    reference (leading dimension n), nothing launches a zero-sized grid, a ray rendered alone equals the same ray in a batch."""
    o, d = workload_rays(max(n, 1), seed=9)
    o, d = o[:n], d[:n]
    pc, pf = O.make_params(4, sigma_bias=0.5, **ARCH), O.make_params(5, sigma_bias=0.5, **ARCH)
    coarse, fine = module_from_params(pc, ARCH), module_from_params(pf, ARCH)
    kw = _kwargs(cn, coarse, fine)
    rgb, disp, acc, depth, ex = cn.render(1, n, None, chunk=4096, rays=(o.to(DEV), d.to(DEV)), retraw=True, **kw)
    assert rgb.shape == (n, 3) and disp.shape == (n,) and acc.shape == (n,) and depth.shape == (n,)
    assert ex["raw"].shape == (n, 192, 4) and ex["rgb0"].shape == (n, 3) and ex["z_std"].shape == (n,)
    loss = (rgb ** 2).sum() + (ex["rgb0"] ** 2).sum()
    loss.backward()
    torch.cuda.synchronize()
    for p in fine.hot_params():
        assert p.grad is not None and torch.isfinite(p.grad).all()
        if n == 0:
            assert float(p.grad.abs().max()) == 0.0
    if n == 129:
        with torch.no_grad():
            one = cn.render(1, 1, None, chunk=4096, rays=(o[128:129].to(DEV), d[128:129].to(DEV)), **kw)
        assert torch.equal(one[0], rgb[128:129].detach()) and torch.equal(one[3], depth[128:129].detach())


def test_stage_functions_accept_empty_batches(cn):
    """raw2outputs / sample_pdf / run_network on zero rays return the reference's empty shapes."""
    raw = torch.zeros(0, 64, 4, device=DEV, requires_grad=True)
    z, d = torch.zeros(0, 64, device=DEV), torch.zeros(0, 3, device=DEV)
    rgb, disp, acc, w, depth = cn.raw2outputs(raw, z, d, 0.0, True)
    assert rgb.shape == (0, 3) and disp.shape == (0,) and acc.shape == (0,) and w.shape == (0, 64) and depth.shape == (0,)
    assert cn.sample_pdf(torch.zeros(0, 63, device=DEV), torch.zeros(0, 62, device=DEV), 128, det=True).shape == (0, 128)
    net = module_from_params(O.make_params(1, **ARCH), ARCH)
    e, _ = cn.get_embedder(10, 0)
    ed, _ = cn.get_embedder(4, 0)
    out = cn.run_network(torch.zeros(0, 64, 3, device=DEV), torch.zeros(0, 3, device=DEV), net, e, ed)
    assert out.shape == (0, 64, 4)
    out.sum().backward()
    assert all(float(p.grad.abs().max()) == 0.0 for p in net.hot_params())
