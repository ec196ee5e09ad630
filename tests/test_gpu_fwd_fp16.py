"""The fp16-operand forward (csrc/mlp_fwd5.cu: one MMA per MAC, two tiles in flight per SM): parity against the fp64 oracle at
the MLP output and at the rendered maps (north-star bar: 1e-4 on rgb / depth / weights), ragged / tiny / multi-tile shapes,
training record + gradients, and determinism."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from util import ARCH, module_from_params, rel_err, workload_rays

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cn():
    import consistentnerf_b200 as m
    return m


def _net(seed, bias=0.3):
    p = O.make_params(seed, sigma_bias=bias, **ARCH)
    net = module_from_params(p, ARCH)
    packed = net.packed_weights()
    packed.refresh({k: v.detach() for k, v in zip(net.spec.param_names(), net.hot_params())})
    return p, net, packed


@pytest.mark.parametrize("n_rays,n_samples", [(1, 1), (3, 64), (2, 128), (40, 192), (257, 33), (1200, 64)])
def test_fp16_forward_vs_fp64_oracle(cn, n_rays, n_samples):
    """1 tile, exactly 2 tiles, ragged last tile, odd tile counts per CTA, more tiles than 2 x 148."""
    p, net, packed = _net(7)
    gen = torch.Generator().manual_seed(n_rays)
    pts = torch.randn(n_rays, n_samples, 3, generator=gen) * 2.0
    vd = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=gen), dim=-1)
    raw = cn.ops.fused_mlp_forward(packed, pts.to(DEV), vd.to(DEV), fwd_terms=1)
    ref = O._query({k: v.double() for k, v in p.items()}, ARCH, pts.double(), vd.double(), 10, 4)
    raw3 = cn.ops.fused_mlp_forward(packed, pts.to(DEV), vd.to(DEV), fwd_terms=3)
    err, err3 = rel_err(raw, ref), rel_err(raw3, ref)
    print(f"fp16 forward rel err vs fp64: {err:.2e} (three-term: {err3:.2e})")
    assert err < 1e-3 and err3 < 2e-5
    again = cn.ops.fused_mlp_forward(packed, pts.to(DEV), vd.to(DEV), fwd_terms=1)
    assert torch.equal(raw, again)                                        # deterministic


def test_fp16_forward_rendered_maps_within_the_parity_bar(cn):
    """render() on 512 rays of workload A with the fp16 forward vs the fp64 oracle: rgb / depth / acc within 1e-4."""
    n = 512
    pc, coarse, _ = _net(0, 0.5)
    pf, fine, _ = _net(1, 0.5)
    e, _ = cn.get_embedder(10, 0)
    ev, _ = cn.get_embedder(4, 0)
    q = lambda i, v, f: cn.run_network(i, v, f, embed_fn=e, embeddirs_fn=ev)
    kw = dict(network_query_fn=q, perturb=0.0, N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse, use_viewdirs=True,
              white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False, near=2.0, far=6.0)
    o, d = workload_rays(n)
    prev = cn.ops.set_forward_precision("fp16")
    try:
        with torch.no_grad():
            rgb, disp, acc, depth, ex = cn.render(1, n, None, chunk=4096, rays=(o.to(DEV), d.to(DEV)), **kw)
    finally:
        cn.ops.set_forward_precision(prev)
    ref = O.render_rays(O.pack_rays(o.double(), d.double(), 2.0, 6.0, True), {k: v.double() for k, v in pc.items()},
                        {k: v.double() for k, v in pf.items()}, ARCH, n_samples=64, n_importance=128, white_bkgd=True)
    errs = {k: rel_err(v, ref[k]) for k, v in (("rgb_map", rgb), ("depth_map", depth), ("acc_map", acc), ("rgb0", ex["rgb0"]), ("depth0", ex["depth0"]))}
    print("fp16 forward, rendered maps vs fp64:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert all(v < 1e-4 for v in errs.values()), errs


def test_fp16_training_forward_record_and_gradients(cn):
    """fwd5 training variant: same raw as its inference variant, hi-only record decodes to the oracle's activations, and the
    fp16 backward on top of it gives gradients close to fp64."""
    from test_gpu_mlp_bwd import GTILE, SLOT_E, SLOT_F, SLOT_H0, SLOT_HV, SLOT_V, TILE, decode, reference_chain
    p, net, packed = _net(17)
    gen = torch.Generator().manual_seed(3)
    n, S = 9, 100                                   # 900 points: 8 tiles, the last one ragged
    pts = torch.randn(n, S, 3, generator=gen) * 1.5
    vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    d_raw = torch.randn(n * S, 4, generator=gen) * 1e-6
    raw, acts = cn.ops.fused_mlp_forward_train(packed, pts.to(DEV), vd.to(DEV), dw_terms=1, fwd_terms=1)
    assert torch.equal(raw, cn.ops.fused_mlp_forward(packed, pts.to(DEV), vd.to(DEV), fwd_terms=1))
    ref = reference_chain(p, pts, vd, d_raw)
    N = n * S
    tol = 2e-3                                      # fp16 storage (2^-11) on top of the fp16-operand forward
    assert rel_err(decode(acts, TILE, SLOT_E, 8, 16384, N, use_lo=False)[:, :63], ref["e"]) < tol
    for l in range(8):
        assert rel_err(decode(acts, TILE, SLOT_H0 + l * 131072, 32, 65536, N, use_lo=False), ref["post"][l]) < tol, l
    assert rel_err(decode(acts, TILE, SLOT_F, 32, 65536, N, use_lo=False), ref["feat"]) < tol
    assert rel_err(decode(acts, TILE, SLOT_V, 4, 16384, N, use_lo=False)[:, :27], ref["ev"]) < tol
    assert rel_err(decode(acts, TILE, SLOT_HV, 16, 32768, N, use_lo=False), ref["hv"]) < tol
    P = {k: v.detach() for k, v in zip(net.spec.param_names(), net.hot_params())}
    grads = cn.ops.fused_mlp_backward(packed, P, acts, d_raw.to(DEV), N, terms=(1, 1))
    a = torch.cat([grads[k].double().cpu().reshape(-1) for k in sorted(grads)])
    b = torch.cat([ref["p64"][k].grad.reshape(-1) for k in sorted(grads)])
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    worst = max(rel_err(grads[k], ref["p64"][k].grad) for k in grads)
    print(f"fp16 forward + fp16 backward: cosine {cos:.8f}, worst per-tensor rel err {worst:.2e}")
    # (the fp16 forward moves a few ReLU masks of near-zero pre-activations, so this is the exact gradient of a slightly different
    # function; with 900 random-sign contributions per parameter that shows up as ~2 % of the heavily cancelled sum)
    assert cos > 0.999 and worst < 0.1


def test_forward_precision_switch(cn):
    assert cn.ops.forward_precision() in cn.ops.FWD_PRECISIONS
    with pytest.raises(ValueError):
        cn.ops.set_forward_precision("bf16")
    # a three-term weight gradient needs the hi/lo record: the training forward then stays three-term whatever the switch says
    p, net, packed = _net(5)
    pts = torch.randn(2, 64, 3, generator=torch.Generator().manual_seed(0)).to(DEV)
    vd = torch.nn.functional.normalize(torch.randn(2, 3, generator=torch.Generator().manual_seed(1)), dim=-1).to(DEV)
    raw_a, _ = cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=3, fwd_terms=1)
    assert torch.equal(raw_a, cn.ops.fused_mlp_forward(packed, pts, vd, fwd_terms=3))
