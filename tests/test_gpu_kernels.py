"""Stage-wise parity of every CUDA kernel (called through the C ABI) against the golden vectors
produced by the unmodified reference and against the CPU oracle on seeded inputs.

Tolerances: integer / index / mask outputs are bit exact; floating point is bounded by the
north-star 1e-4 relative bar (most stages are far tighter and say so)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import nerf_oracle as O
from util import ARCH, SMALL, assert_close, module_from_params, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cn():
    import consistentnerf_b200 as m
    return m


# ------------------------------------------------------------------------------------------
# rays / K1
# ------------------------------------------------------------------------------------------
def test_image_rays_and_ndc(cn):
    g = load_golden("rays")
    H, W = (int(v) for v in g["hw"])
    ro, rd = cn.get_rays(H, W, g["K"], t(g["c2w"], device=DEV))
    assert_close(ro, g["rays_o"], 0, 0)
    assert_close(rd, g["rays_d"], 1e-6, 1e-7)
    no, nd = cn.ndc_rays(H, W, float(g["focal"]), 1.0, t(g["rays_o"], device=DEV), t(g["rays_d"], device=DEV))
    assert_close(no, g["ndc_o"], 2e-6, 1e-6)
    assert_close(nd, g["ndc_d"], 2e-6, 1e-6)


def test_pack_rays_viewdirs(cn):
    o = torch.randn(1000, 3, generator=torch.Generator().manual_seed(1))
    d = torch.randn(1000, 3, generator=torch.Generator().manual_seed(2))
    ref = O.pack_rays(o, d, 2.0, 6.0, True)
    got = cn.ops.pack_rays(o.to(DEV), d.to(DEV), 2.0, 6.0, True)
    assert got.shape == (1000, 11)
    assert_close(got, ref, 1e-6, 1e-7)


@pytest.mark.parametrize("lindisp", [False, True])
@pytest.mark.parametrize("perturb", [False, True])
def test_stratified_z_bit_exact(cn, lindisp, perturb):
    n, S = 333, 64
    gen = torch.Generator().manual_seed(3)
    o, d = torch.randn(n, 3, generator=gen), torch.randn(n, 3, generator=gen)
    near = 0.5 + torch.rand(n, 1, generator=gen)
    far = near + 1.0 + 4 * torch.rand(n, 1, generator=gen)
    rays = torch.cat([o, d, near, far], -1)
    t_vals = torch.linspace(0.0, 1.0, S)
    t_rand = torch.rand(n, S, generator=gen) if perturb else None
    z_ref = O.stratified_z(near, far, S, lindisp, t_rand, t_vals=t_vals)
    z, pts = cn.ops.stratified(rays.to(DEV), t_vals.to(DEV), t_rand.to(DEV) if perturb else None, lindisp)
    assert torch.equal(z.cpu(), z_ref)
    assert torch.equal(pts.cpu(), O.ray_points(o, d, z_ref))
    assert torch.equal(cn.ops.ray_points(rays.to(DEV), z).cpu(), O.ray_points(o, d, z_ref))


# ------------------------------------------------------------------------------------------
# K2 encoding
# ------------------------------------------------------------------------------------------
def test_posenc_golden(cn):
    g = load_golden("embed")
    x = t(g["x"], device=DEV)
    e10, d10 = cn.get_embedder(10, 0)
    e4, d4 = cn.get_embedder(4, 0)
    assert (d10, d4) == (63, 27)
    # sin/cos at |arg| up to ~3e3 rad: full-range device sincos vs glibc, a few ulp
    assert_close(e10(x), g["e10"], 0, 2e-6)
    assert_close(e4(x), g["e4"], 0, 1e-6)
    ident, d = cn.get_embedder(10, -1)
    assert d == 3 and torch.equal(ident(x), x)


def test_posenc_large_arguments(cn):
    x = (torch.rand(4096, 3, generator=torch.Generator().manual_seed(5)) * 12 - 6)
    ref = O.posenc(x.double(), 10)
    got = cn.ops.posenc(x.to(DEV), 10)
    assert float((got.cpu().double() - ref).abs().max()) < 2e-6


# ------------------------------------------------------------------------------------------
# K3 generic fp32 layers
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,arch", [("mlp_viewdirs", ARCH), ("mlp_small_noview", SMALL)])
def test_layerwise_mlp_golden(cn, tag, arch):
    g = load_golden(tag)
    p = O.make_params(int(g["seed"]), **arch)
    net = module_from_params(p, arch)
    with torch.no_grad():
        y = net(t(g["x"], device=DEV))
    assert_close(y, g["y"], 2e-5, 2e-6)


def test_layerwise_mlp_backward_matches_autograd(cn):
    arch = SMALL
    p = O.make_params(11, **arch)
    net = module_from_params(p, arch)
    x = torch.randn(517, 63, generator=torch.Generator().manual_seed(4))
    gy = torch.randn(517, 5, generator=torch.Generator().manual_seed(6))
    net(x.to(DEV)).backward(gy.to(DEV))
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    O.mlp_forward(p64, x.double(), **arch).backward(gy.double())
    for name, prm in net.named_parameters():
        if p64[name].grad is None:
            continue
        assert rel_err(prm.grad, p64[name].grad) < 2e-5, name


# ------------------------------------------------------------------------------------------
# K2+K3 fused tcgen05 path
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k", [(128, 64), (256, 32), (16, 16), (48, 128)])
def test_umma_building_blocks(cn, n, k):
    gen = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=gen)
    b = torch.randn(n, k, generator=gen) * 0.1
    d = cn.ops.umma_selftest(a.to(DEV), b.to(DEV))
    ref = a.double() @ b.double().t()
    assert rel_err(d, ref) < 2e-6


@pytest.mark.parametrize("n,k", [(128, 64), (256, 128), (16, 16), (48, 128), (128, 256)])
def test_umma_a_operand_in_tensor_memory(cn, n, k):
    gen = torch.Generator().manual_seed(n * 1000 + k + 1)
    a = torch.randn(128, k, generator=gen)
    b = torch.randn(n, k, generator=gen) * 0.1
    d = cn.ops.umma_selftest(a.to(DEV), b.to(DEV), a_in_tmem=True)
    assert rel_err(d, a.double() @ b.double().t()) < 2e-6


def _needs_experiments():
    from consistentnerf_b200 import _lib
    if not hasattr(_lib.load(), "cnerf_umma_selftest_pair"):
        pytest.skip("library built without csrc/experiments (python -m consistentnerf_b200.build --experiments)")


@pytest.mark.parametrize("n,k", [(256, 16), (256, 128), (128, 64), (32, 32)])
def test_umma_cta_pair(cn, n, k):
    """One M=256 tcgen05.mma.cta_group::2 stream for two CTAs: each holds 128 rows of A/D and n/2 rows of B."""
    _needs_experiments()
    gen = torch.Generator().manual_seed(n * 1000 + k + 2)
    a = torch.randn(256, k, generator=gen)
    b = torch.randn(n, k, generator=gen) * 0.1
    d = cn.ops.umma_selftest(a.to(DEV), b.to(DEV))
    assert rel_err(d, a.double() @ b.double().t()) < 2e-6


@pytest.mark.parametrize("n_rays,n_samples", [(1, 1), (3, 64), (40, 192), (257, 33)])
def test_fused_mlp_forward(cn, n_rays, n_samples):
    p = O.make_params(7, sigma_bias=0.3, **ARCH)
    net = module_from_params(p, ARCH)
    gen = torch.Generator().manual_seed(n_rays)
    pts = torch.randn(n_rays, n_samples, 3, generator=gen) * 2.0
    vd = torch.randn(n_rays, 3, generator=gen)
    vd = vd / vd.norm(dim=-1, keepdim=True)
    packed = net.packed_weights()
    packed.refresh({k: v for k, v in zip(net.spec.param_names(), net.hot_params())})
    raw = cn.ops.fused_mlp_forward(packed, pts.to(DEV), vd.to(DEV))
    p64 = {k: v.double() for k, v in p.items()}
    ref = O._query(p64, ARCH, pts.double(), vd.double(), 10, 4)
    ref32 = O._query(p, ARCH, pts, vd, 10, 4)
    err, err32 = rel_err(raw, ref), rel_err(ref32, ref)
    print(f"fused MLP rel err vs fp64: {err:.2e} (fp32 oracle: {err32:.2e})")
    assert err < 2e-5
    assert_close(raw, ref, 1e-4, 2e-5)


def test_fused_mlp_repacks_after_inplace_update(cn):
    p = O.make_params(8, **ARCH)
    net = module_from_params(p, ARCH)
    e, _ = cn.get_embedder(10)
    ev, _ = cn.get_embedder(4)
    pts = torch.randn(5, 64, 3, generator=torch.Generator().manual_seed(0)).to(DEV)
    vd = torch.nn.functional.normalize(torch.randn(5, 3, generator=torch.Generator().manual_seed(1)), dim=-1).to(DEV)
    with torch.no_grad():
        a = cn.run_network(pts, vd, net, e, ev)
        net.alpha_linear.bias.add_(1.0)              # what optimizer.step() does: in-place
        b = cn.run_network(pts, vd, net, e, ev)
    assert_close(b[..., 3] - a[..., 3], torch.ones(5, 64), 0, 1e-5)
    assert torch.equal(a[..., :3], b[..., :3])


# ------------------------------------------------------------------------------------------
# K4 compositing
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,wb", [("raw2outputs_wb0", False), ("raw2outputs_wb1", True), ("raw2outputs_noise", True)])
def test_composite_golden(cn, tag, wb):
    g = load_golden(tag)
    noise = t(g["noise"], device=DEV) if "noise" in g else None
    rgb, disp, acc, w, depth = cn.ops.CompositeFn.apply(t(g["raw"], device=DEV), t(g["z"], device=DEV),
                                                        t(g["rays_d"], device=DEV), noise, wb)
    for name, got in (("rgb", rgb), ("acc", acc), ("weights", w), ("depth", depth)):
        assert_close(got, g[name], 2e-5, 2e-6, name)
    assert_close(disp, g["disp"], 1e-4, 1e-6, "disp")


def test_composite_edge_cases(cn):
    # all-zero density -> weights 0, acc 0, disp NaN (0/0), white background = 1
    raw = torch.zeros(4, 16, 4)
    raw[..., 3] = -1.0
    z = torch.linspace(2, 6, 16).expand(4, 16).contiguous()
    d = torch.tensor([[0.0, 0.0, -1.0]]).expand(4, 3).contiguous()
    ref = O.composite(raw, z, d, None, True)
    rgb, disp, acc, w, depth = cn.ops.CompositeFn.apply(raw.to(DEV), z.to(DEV), d.to(DEV), None, True)
    assert torch.equal(w.cpu(), ref["weights"]) and torch.equal(acc.cpu(), ref["acc"])
    assert torch.isnan(disp).all() and torch.isnan(ref["disp"]).all()
    assert_close(rgb, ref["rgb"], 0, 0)
    # two samples (the reference needs S >= 2), ragged S (not a multiple of 32), opaque first sample
    for S in (2, 33, 95):
        gen = torch.Generator().manual_seed(S)
        raw = torch.randn(7, S, 4, generator=gen) * 3
        raw[0, 0, 3] = 50.0
        z = torch.sort(2 + 4 * torch.rand(7, S, generator=gen), -1).values
        d = torch.randn(7, 3, generator=gen)
        ref = O.composite(raw.double(), z.double(), d.double(), None, False)
        out = cn.ops.CompositeFn.apply(raw.to(DEV), z.to(DEV), d.to(DEV), None, False)
        for got, key in zip(out, ("rgb", "disp", "acc", "weights", "depth")):
            assert_close(got, ref[key].float(), 1e-4, 1e-6, f"S={S} {key}")


def test_composite_backward(cn):
    gen = torch.Generator().manual_seed(9)
    n, S = 50, 70
    raw = torch.randn(n, S, 4, generator=gen)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=gen), -1).values
    d = torch.randn(n, 3, generator=gen)
    noise = 0.3 * torch.randn(n, S, generator=gen)
    gs = [torch.randn(n, 3, generator=gen), torch.randn(n, generator=gen), torch.randn(n, generator=gen),
          torch.randn(n, S, generator=gen), torch.randn(n, generator=gen)]
    r64 = raw.double().requires_grad_(True)
    c = O.composite(r64, z.double(), d.double(), noise.double(), True)
    loss = sum((c[k] * g.double()).sum() for k, g in zip(("rgb", "disp", "acc", "weights", "depth"), gs))
    loss.backward()
    rg = raw.to(DEV).requires_grad_(True)
    out = cn.ops.CompositeFn.apply(rg, z.to(DEV), d.to(DEV), noise.to(DEV), True)
    sum((o * g.to(DEV)).sum() for o, g in zip(out, gs)).backward()
    assert rel_err(rg.grad, r64.grad) < 1e-4


# ------------------------------------------------------------------------------------------
# K5 hierarchical sampling
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["det", "rnd"])
def test_sample_pdf_golden_bit_exact(cn, mode):
    g = load_golden("sample_pdf")
    s, dbg = cn.ops.sample_pdf(t(g["bins"], device=DEV), t(g["weights"], device=DEV), t(g["u_" + mode], device=DEV),
                               128, debug=True)
    assert torch.equal(dbg["cdf"].cpu(), t(g["cdf_" + mode]))
    inds = t(g["inds_" + mode])
    assert torch.equal(dbg["below"].cpu().long(), torch.clamp(inds - 1, min=0))
    assert torch.equal(dbg["above"].cpu().long(), torch.clamp(inds, max=62))
    assert torch.equal(s.cpu(), t(g["samples_" + mode]))


def test_sample_pdf_det_linspace_and_module_api(cn):
    g = load_golden("sample_pdf")
    bins, w = t(g["bins"], device=DEV), t(g["weights"], device=DEV)
    s = cn.sample_pdf(bins, w, 128, det=True)
    assert torch.equal(s.cpu(), t(g["samples_det"]))
    # pytest hook + det: the reference takes u from np.linspace (float64 -> float32), NP/run_nerf_helpers.py:222-225
    u_np = torch.tensor(np.broadcast_to(np.linspace(0.0, 1.0, 128), (48, 128)).copy(), dtype=torch.float32)
    assert torch.equal(cn.sample_pdf(bins, w, 128, det=True, pytest=True).cpu(),
                       O.sample_pdf(t(g["bins"]), t(g["weights"]), u_np))


def test_sample_pdf_degenerate_weights(cn):
    # all-zero weights (uniform pdf), one-hot weights, ragged bin counts
    for B in (2, 5, 63, 130):
        gen = torch.Generator().manual_seed(B)
        bins = torch.sort(torch.rand(9, B, generator=gen), -1).values
        w = torch.rand(9, B - 1, generator=gen)
        w[0] = 0.0
        w[1] = 0.0
        w[1, (B - 1) // 2] = 1.0
        u = torch.rand(9, 37, generator=gen)
        u[2, 0], u[2, 1] = 0.0, 0.99999994
        ref, dbg = O.sample_pdf(bins, w, u, return_debug=True)
        got, gd = cn.ops.sample_pdf(bins.to(DEV), w.to(DEV), u.to(DEV), 37, debug=True)
        assert torch.equal(gd["cdf"].cpu(), dbg["cdf"]), B
        assert torch.equal(gd["below"].cpu().long(), dbg["below"]) and torch.equal(gd["above"].cpu().long(), dbg["above"])
        assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("det", [True, False])
def test_sample_fine_fused(cn, det):
    n, S, M = 211, 64, 128
    gen = torch.Generator().manual_seed(12)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=gen), -1).values
    w = torch.rand(n, S, generator=gen) ** 4
    u = None if det else torch.rand(n, M, generator=gen)
    mid = 0.5 * (z[:, 1:] + z[:, :-1])
    uu = torch.linspace(0.0, 1.0, M).expand(n, M) if det else u
    zs_ref = O.sample_pdf(mid, w[:, 1:-1], uu)
    zf_ref = O.merge_sorted(z, zs_ref)
    zs, zf, zstd = cn.ops.sample_fine(z.to(DEV), w.to(DEV), None if det else u.to(DEV), M)
    assert torch.equal(zs.cpu(), zs_ref)
    assert torch.equal(zf.cpu(), zf_ref)                       # sortedness + exact multiset
    assert_close(zstd, torch.std(zs_ref.double(), -1, unbiased=False).float(), 1e-6, 1e-7)


# ------------------------------------------------------------------------------------------
# K6 cross-view geometry
# ------------------------------------------------------------------------------------------
def test_project_gather_golden(cn):
    g = load_golden("crossview")
    H, W = g["depth_ref"].shape
    res = cn.ops.project_gather(t(g["pts_w"], device=DEV), t(g["w2c_ref"]), t(g["K"]), H, W,
                                img=t(g["img_ref"], device=DEV), depth=t(g["depth_ref"], device=DEV), c2w=t(g["c2w_ref"]))
    inb = res["mask"].bool().cpu()
    assert np.array_equal(res["px"].cpu().numpy(), g["label_x"]) and np.array_equal(res["py"].cpu().numpy(), g["label_y"])
    assert np.array_equal(inb.numpy(), g["inb"])
    assert_close(res["cam"], g["cam"], 2e-6, 2e-6)
    assert np.array_equal(res["rgb"].cpu()[inb].numpy(), g["rgb_ref"])
    assert np.array_equal(res["depth"].cpu()[inb].numpy(), g["dep_ref"])
    assert_close(res["rays_o"].cpu()[inb], g["ref_rays_o"], 1e-6, 1e-7)
    assert_close(res["rays_d"].cpu()[inb], g["ref_rays_d"], 1e-6, 1e-6)


def test_reference_surface_get_ref_rays_and_test_label(cn):
    g = load_golden("crossview")
    dev = DEV
    w2c, c2w, K = (t(g[k], device=dev)[None] for k in ("w2c_ref", "c2w_ref", "K"))
    pts = t(g["pts_w"], device=dev)[None, :, None, :]
    img = t(g["img_ref"], device=dev)[None]
    dep = t(g["depth_ref"], device=dev)[None]
    rgb_ref, depth_ref, cam, ro, rd, mask = cn.get_ref_rays(w2c, c2w, K, pts, img, dep)
    assert rgb_ref.shape == (1, 3, 661) and depth_ref.shape == (1, 1, 661) and mask.shape == (1, 768)
    assert np.array_equal(rgb_ref[0].t().cpu().numpy(), g["rgb_ref"])
    assert np.array_equal(depth_ref.reshape(-1).cpu().numpy(), g["dep_ref"])
    y, x, m, zc = cn.get_test_label(w2c, c2w, K, pts, img)
    assert np.array_equal(y[0].cpu().numpy(), g["label_y"]) and np.array_equal(x[0].cpu().numpy(), g["label_x"])
    assert np.array_equal(m[0].cpu().numpy(), g["label_mask"])
    assert_close(zc[0], g["label_z"], 2e-6, 2e-6)


def test_hard_mask_golden_and_chunk_semantics(cn):
    g = load_golden("crossview")
    args = (t(g["rays_o"], device=DEV), t(g["rays_d"], device=DEV), t(g["depth_tgt"], device=DEV).reshape(-1),
            t(g["w2c_ref"]), t(g["K"]), t(g["depth_ref"], device=DEV))
    m = cn.ops.hard_mask_pair(*args, thr0=float(g["thr0"]), chunk=int(g["chunk"]))
    assert np.array_equal(m.bool().cpu().numpy(), g["hard_mask"])
    cpu = [a.cpu() if a.is_cuda else a for a in args]
    for chunk, thr0 in ((100, 1e-4), (5120, 1e-3), (37, 0.1)):
        ref = O.hard_mask_pair(*cpu, thr0=thr0, chunk=chunk)
        got = cn.ops.hard_mask_pair(*args, thr0=thr0, chunk=chunk)
        assert torch.equal(got.bool().cpu(), ref), (chunk, thr0)
    # OR accumulation over reference views
    acc = torch.ones(768, device=DEV, dtype=torch.uint8)
    cn.ops.hard_mask_pair(*args, thr0=0.1, chunk=5120, mask=acc)
    assert bool(acc.all())


# ------------------------------------------------------------------------------------------
# K7 masked losses
# ------------------------------------------------------------------------------------------
def test_masked_losses_golden(cn):
    g = load_golden("masked_loss")
    rgb, tgt, m = (t(g[k], device=DEV) for k in ("rgb", "tgt", "mask"))
    far, coef = float(g["far"]), float(g["coef"])
    li = cn.masked_img_loss(rgb, tgt, m, coef)
    assert_close(li, g["img_loss"], 1e-6, 0)
    assert_close(cn.mse2psnr(li), g["psnr"].reshape(()), 1e-6, 0)
    d, q = t(g["depth"], device=DEV), t(g["depth_prior"], device=DEV)
    assert_close(cn.masked_depth_loss(d, q, m, far), g["depth_loss_masked_only"], 1e-6, 0)
    assert_close(cn.masked_depth_loss(d, q, m, far, coef, include_unmasked=True), g["depth_loss_both"], 1e-6, 0)
    assert_close(cn.img2mse(rgb, tgt), torch.mean((t(g["rgb"]) - t(g["tgt"])) ** 2), 1e-6, 0)


def test_masked_loss_all_ones_mask_and_backward(cn):
    gen = torch.Generator().manual_seed(21)
    n = 4099
    rgb, tgt = torch.rand(n, 3, generator=gen), torch.rand(n, 3, generator=gen)
    for mask in (torch.ones(n, 1), (torch.rand(n, 1, generator=gen) > 0.4).float()):
        p64 = rgb.double().requires_grad_(True)
        ref = O.masked_mse(p64, tgt.double(), mask.double(), 0.2, n_ref=n)
        ref.backward()
        p = rgb.to(DEV).requires_grad_(True)
        got = cn.masked_img_loss(p, tgt.to(DEV), mask.to(DEV), 0.2)
        got.backward()
        assert_close(got, ref.float(), 1e-6, 0)
        assert rel_err(p.grad, p64.grad) < 1e-5


# ------------------------------------------------------------------------------------------
# opt-in kernel generations (selected once per process by an environment variable -> run in a subprocess)
# ------------------------------------------------------------------------------------------
_IMPL_SCRIPT = r"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "tests")); sys.path.insert(0, os.getcwd())
import consistentnerf_b200 as cn
from oracle import nerf_oracle as O
from util import ARCH, module_from_params, rel_err
p = O.make_params(7, sigma_bias=0.3, **ARCH)
net = module_from_params(p, ARCH)
packed = net.packed_weights(); packed.refresh(dict(zip(net.spec.param_names(), [q.detach() for q in net.hot_params()])))
worst = 0.0
for n, S in ((1, 1), (3, 64), (40, 192), (257, 33), (700, 64)):
    g = torch.Generator().manual_seed(n * 100 + S)
    pts = torch.randn(n, S, 3, generator=g) * 1.5
    vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    raw = cn.ops.fused_mlp_forward(packed, pts.cuda(), vd.cuda())
    ref = O._query({k: v.double() for k, v in p.items()}, ARCH, pts.double(), vd.double(), 10, 4)
    worst = max(worst, rel_err(raw, ref))
print("WORST", worst)
assert worst < 2e-5, worst
"""


@pytest.mark.parametrize("env", [{"CNERF_MLP_IMPL": "4"}])
def test_opt_in_forward_kernels_match_the_oracle(env):
    """CNERF_MLP_IMPL=4: CTA-pair ping-pong kernel (experiments/mlp_fwd4.cu)."""
    _needs_experiments()
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", _IMPL_SCRIPT], cwd=root, env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
