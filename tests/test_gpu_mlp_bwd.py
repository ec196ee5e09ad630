"""Stage-wise parity of the tensor-core MLP backward (K3b): activation record of the training forward, the
data-gradient chain's G tiles, and the parameter gradients, each against an fp64 restatement
(NP/run_nerf_helpers.py:107-130 differentiated by torch autograd)."""
import math

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from util import ARCH, module_from_params, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

TILE = 1310720 + 34816          # operand slots + the ReLU sign-bit slot M
SLOT_E, SLOT_H0, SLOT_F, SLOT_V, SLOT_HV = 0, 32768, 32768 + 8 * 131072, 32768 + 9 * 131072, 32768 + 9 * 131072 + 32768
SLOT_M = SLOT_HV + 65536
GTILE = 65536 + 9 * 131072


def g_slot(l):
    return 0 if l == 9 else 65536 + (8 - l) * 131072


def decode(rec_u8, tile_bytes, slot, kgroups, lo_off, n_points, use_lo=True):
    """[n_points, 8*kgroups] float64 = hi (+ lo) from the fp16 UMMA tiles of every 128-point record."""
    rec = rec_u8.cpu().numpy().reshape(-1, tile_bytes)
    out = []
    for t in range(rec.shape[0]):
        hi = rec[t, slot:slot + kgroups * 2048].view(np.float16).reshape(kgroups, 128, 8)
        lo = rec[t, slot + lo_off:slot + lo_off + kgroups * 2048].view(np.float16).reshape(kgroups, 128, 8)
        v = hi.astype(np.float64) + (lo.astype(np.float64) if use_lo else 0.0)
        out.append(v.transpose(1, 0, 2).reshape(128, kgroups * 8))
    return torch.from_numpy(np.concatenate(out, 0)[:n_points])


def reference_chain(p, pts, vd, d_raw):
    """fp64 forward keeping every pre-activation as a leaf-like tensor with retained gradient."""
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    n, S = pts.shape[:2]
    e = O.posenc(pts.double().reshape(-1, 3), 10)
    ev = O.posenc(vd.double()[:, None, :].expand(n, S, 3).reshape(-1, 3), 4)
    pre, post = [], []
    h = e
    for i in range(8):
        z = h @ p64[f"pts_linears.{i}.weight"].t() + p64[f"pts_linears.{i}.bias"]
        z.retain_grad(); pre.append(z)
        h = torch.relu(z); post.append(h)
        if i == 4:
            h = torch.cat([e, h], -1)
    sigma = h @ p64["alpha_linear.weight"].t() + p64["alpha_linear.bias"]
    feat = h @ p64["feature_linear.weight"].t() + p64["feature_linear.bias"]
    feat.retain_grad()
    zv = torch.cat([feat, ev], -1) @ p64["views_linears.0.weight"].t() + p64["views_linears.0.bias"]
    zv.retain_grad()
    hv = torch.relu(zv)
    rgb = hv @ p64["rgb_linear.weight"].t() + p64["rgb_linear.bias"]
    raw = torch.cat([rgb, sigma], -1)
    raw.backward(d_raw.double())
    return dict(p64=p64, e=e, ev=ev, pre=pre, post=post, feat=feat, zv=zv, hv=hv, raw=raw)


@pytest.fixture(scope="module")
def setup():
    import consistentnerf_b200 as cn
    p = O.make_params(17, sigma_bias=0.3, **ARCH)
    net = module_from_params(p, ARCH)
    gen = torch.Generator().manual_seed(3)
    n, S = 5, 60                                   # 300 points: two full tiles + a ragged one
    pts = torch.randn(n, S, 3, generator=gen) * 1.5
    vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    d_raw = torch.randn(n * S, 4, generator=gen) * 1e-6      # realistic tiny loss gradients
    packed = net.packed_weights()
    P = {k: v.detach() for k, v in zip(net.spec.param_names(), net.hot_params())}
    packed.refresh(P)
    raw, acts = cn.ops.fused_mlp_forward_train(packed, pts.to(DEV), vd.to(DEV))
    ref = reference_chain(p, pts, vd, d_raw)
    return dict(cn=cn, p=p, net=net, packed=packed, P=P, pts=pts, vd=vd, d_raw=d_raw, raw=raw, acts=acts, ref=ref, n=n * S)


def test_training_forward_equals_inference_forward(setup):
    cn = setup["cn"]
    raw2 = cn.ops.fused_mlp_forward(setup["packed"], setup["pts"].to(DEV), setup["vd"].to(DEV))
    assert torch.equal(raw2, setup["raw"])
    assert rel_err(setup["raw"].reshape(-1, 4), setup["ref"]["raw"]) < 2e-5


def test_activation_record(setup):
    acts, ref, n = setup["acts"], setup["ref"], setup["n"]
    assert acts.numel() == 3 * TILE
    E = decode(acts, TILE, SLOT_E, 8, 16384, n)
    assert rel_err(E[:, :63], ref["e"]) < 2e-6
    assert float((E[:, 63] - 1.0).abs().max()) == 0.0        # the constant column that carries the biases through the MMAs
    for l in range(8):
        H = decode(acts, TILE, SLOT_H0 + l * 131072, 32, 65536, n)
        assert rel_err(H, ref["post"][l]) < 5e-6, l
    F = decode(acts, TILE, SLOT_F, 32, 65536, n)
    assert rel_err(F, ref["feat"]) < 5e-6
    V = decode(acts, TILE, SLOT_V, 4, 16384, n)
    assert rel_err(V[:, :27], ref["ev"]) < 2e-6
    HV = decode(acts, TILE, SLOT_HV, 16, 32768, n)
    assert rel_err(HV, ref["hv"]) < 5e-6
    # ReLU sign bits (slot M), bit j of a byte = [feature 8 kg + j > 0]: H_l bytes at l * 4096 + p * 1024 + row * 8 + kb for
    # k-group kb * 4 + p; views-layer bytes at 32768 + row * 16 + kg (what the data-gradient chain masks with).
    rec = acts.cpu().numpy().reshape(-1, TILE)

    def unpack(b):                                   # [128 rows, kgroups] bytes -> [128, 8 * kgroups] bools
        return np.unpackbits(b[:, :, None], axis=2, bitorder="little").reshape(128, -1)

    def hidden_bits(l):
        out = []
        for t_ in range(rec.shape[0]):
            b = rec[t_, SLOT_M + l * 4096:SLOT_M + (l + 1) * 4096].reshape(4, 128, 8)              # [p][row][kb]
            out.append(unpack(b.transpose(1, 2, 0).reshape(128, 32)))                                 # k-group = kb * 4 + p
        return torch.from_numpy(np.concatenate(out, 0)[:n].astype(bool))

    def same_sign(bits_, pre):          # the bits are the signs of the kernel's own fp32 pre-activations: compare away from zero
        pre = pre.detach()
        decided = pre.abs() > 1e-5 * pre.abs().max()
        assert float(decided.float().mean()) > 0.999
        assert torch.equal(bits_[decided], (pre > 0)[decided])

    for l in range(8):
        same_sign(hidden_bits(l), ref["pre"][l])
    hvb = np.concatenate([unpack(rec[t_, SLOT_M + 32768:SLOT_M + 32768 + 2048].reshape(128, 16)) for t_ in range(rec.shape[0])], 0)
    same_sign(torch.from_numpy(hvb[:n].astype(bool)), ref["zv"])


def test_gradient_chain_and_parameter_gradients(setup):
    cn, ref, n = setup["cn"], setup["ref"], setup["n"]
    grads, rec = cn.ops.fused_mlp_backward(setup["packed"], setup["P"], setup["acts"], setup["d_raw"].to(DEV), n,
                                           return_record=True)
    torch.cuda.synchronize()
    amax = float(setup["d_raw"].abs().max())
    scale = 2.0 ** math.floor(math.log2(256.0 / amax))
    errs = {}
    G9 = decode(rec, GTILE, g_slot(9), 16, 32768, n) / scale
    errs["G9"] = rel_err(G9, ref["zv"].grad)
    G8 = decode(rec, GTILE, g_slot(8), 32, 65536, n) / scale
    errs["G8"] = rel_err(G8, ref["feat"].grad)
    for l in range(7, -1, -1):
        G = decode(rec, GTILE, g_slot(l), 32, 65536, n) / scale
        errs[f"G{l}"] = rel_err(G, ref["pre"][l].grad)
    print("chain:", {k: f"{v:.1e}" for k, v in errs.items()})
    perr = {k: rel_err(grads[k], ref["p64"][k].grad) for k in grads}
    print("params:", {k: f"{v:.1e}" for k, v in perr.items()})
    assert all(v < 2e-5 for v in errs.values()), errs
    assert all(v < 2e-5 for v in perr.values()), perr


# (chain_terms, dw_terms) -> tolerance on the chain's G tiles / on the parameter gradients (max-scale relative).  One-term
# products round each operand to fp16 (2^-11 per element, unbiased); over the 300 points of this fixture that leaves ~1e-3.
REDUCED = {"dw16": ((3, 1), 6e-4, 3e-3), "fp16": ((1, 1), 4e-3, 6e-3)}


@pytest.mark.parametrize("mode", sorted(REDUCED))
def test_reduced_precision_backward_modes(setup, mode):
    """The fp16-operand backward modes (include/cnerf.h K3b): hi-only records, one MMA per MAC."""
    cn, ref, n = setup["cn"], setup["ref"], setup["n"]
    terms, tol_g, tol_p = REDUCED[mode]
    raw, acts = cn.ops.fused_mlp_forward_train(setup["packed"], setup["pts"].to(DEV), setup["vd"].to(DEV), dw_terms=terms[1])
    assert torch.equal(raw, setup["raw"])                      # the forward result does not depend on the record format
    grads, rec = cn.ops.fused_mlp_backward(setup["packed"], setup["P"], acts, setup["d_raw"].to(DEV), n, return_record=True,
                                           terms=terms)
    torch.cuda.synchronize()
    amax = float(setup["d_raw"].abs().max())
    scale = 2.0 ** math.floor(math.log2(256.0 / amax))
    errs = {"G9": rel_err(decode(rec, GTILE, g_slot(9), 16, 32768, n, use_lo=False) / scale, ref["zv"].grad),
            "G8": rel_err(decode(rec, GTILE, g_slot(8), 32, 65536, n, use_lo=False) / scale, ref["feat"].grad)}
    for l in range(7, -1, -1):
        errs[f"G{l}"] = rel_err(decode(rec, GTILE, g_slot(l), 32, 65536, n, use_lo=False) / scale, ref["pre"][l].grad)
    perr = {k: rel_err(grads[k], ref["p64"][k].grad) for k in grads}
    print(mode, "chain:", {k: f"{v:.1e}" for k, v in errs.items()})
    print(mode, "params:", {k: f"{v:.1e}" for k, v in perr.items()})
    assert all(v < tol_g for v in errs.values()), errs
    assert all(v < tol_p for v in perr.values()), perr
    # direction of the full gradient: what a training step consumes
    a = torch.cat([grads[k].double().cpu().reshape(-1) for k in sorted(grads)])
    b = torch.cat([ref["p64"][k].grad.reshape(-1) for k in sorted(grads)])
    cos = float((a * b).sum() / (a.norm() * b.norm()))
    assert cos > 1.0 - 1e-5, cos


def test_grad_precision_switch_reaches_the_autograd_path(setup):
    cn, net = setup["cn"], setup["net"]
    e, _ = cn.get_embedder(10)
    ev, _ = cn.get_embedder(4)
    pts, vd = setup["pts"].to(DEV), setup["vd"].to(DEV)
    g = torch.randn(5, 60, 4, generator=torch.Generator().manual_seed(9)).to(DEV) * 1e-5
    out = {}
    prev = cn.ops.grad_precision()
    try:
        for mode in ("split", "dw16", "fp16"):
            cn.ops.set_grad_precision(mode)
            net.zero_grad()
            cn.run_network(pts, vd, net, e, ev).backward(g)
            out[mode] = {k: v.grad.clone() for k, v in net.named_parameters() if v.grad is not None}
    finally:
        cn.ops.set_grad_precision(prev)
    for mode in ("dw16", "fp16"):
        worst = max(rel_err(out[mode][k], out["split"][k]) for k in out["split"])
        assert 0.0 < worst < 6e-3, (mode, worst)                 # really a different arithmetic, and close
    with pytest.raises(ValueError):
        cn.ops.set_grad_precision("bf16")


def test_inplace_accumulation_is_opt_in(setup):
    """parameters that merely own a .grad get their gradients through autograd (torch.autograd.grad works,
    AccumulateGrad hooks fire); only FlatGrads-marked parameters take the in-place path."""
    cn, net = setup["cn"], setup["net"]
    e, _ = cn.get_embedder(10)
    ev, _ = cn.get_embedder(4)
    pts, vd = setup["pts"].to(DEV), setup["vd"].to(DEV)
    for p in net.parameters():
        p.grad = torch.zeros_like(p)                             # every parameter owns a .grad, none is marked
    params = net.hot_params()
    out = cn.run_network(pts, vd, net, e, ev)
    got = torch.autograd.grad(out.sum() * 1e-6, params)
    assert all(g is not None and float(g.abs().max()) > 0 for g in got)
    assert all(float(p.grad.abs().max()) == 0.0 for p in params)          # .grad untouched by autograd.grad
    fired = []
    h = params[0].register_post_accumulate_grad_hook(lambda p: fired.append(1))
    (cn.run_network(pts, vd, net, e, ev).sum() * 1e-6).backward()
    h.remove()
    assert fired == [1]
    from consistentnerf_b200.distributed import FlatGrads
    ref = [p.grad.clone() for p in params]
    flat = FlatGrads(params)
    flat.zero_()
    (cn.run_network(pts, vd, net, e, ev).sum() * 1e-6).backward()
    for p, r in zip(params, ref):
        assert p.grad.data_ptr() >= flat.flat.data_ptr() and rel_err(p.grad, r) < 1e-6
    for p in net.parameters():
        p.grad = None
        if hasattr(p, "_cnerf_accumulate_in_place"):
            del p._cnerf_accumulate_in_place


def test_autograd_path_matches_cuda_core_backward(setup):
    """render-level: the tensor-core backward and the fp32 CUDA-core backward give the same parameter gradients."""
    cn, net = setup["cn"], setup["net"]
    e, _ = cn.get_embedder(10)
    ev, _ = cn.get_embedder(4)
    pts, vd = setup["pts"].to(DEV), setup["vd"].to(DEV)
    g = torch.randn(5, 60, 4, generator=torch.Generator().manual_seed(9)).to(DEV) * 1e-5
    out = {}
    for mode in ("tc", "simt"):
        cn.ops.MLP_BWD = mode
        net.zero_grad()
        cn.run_network(pts, vd, net, e, ev).backward(g)
        out[mode] = {k: v.grad.clone() for k, v in net.named_parameters() if v.grad is not None}
    cn.ops.MLP_BWD = "tc"
    for k in out["tc"]:
        assert rel_err(out["tc"][k], out["simt"][k]) < 2e-5, k


def test_two_tile_kernels_many_tiles_per_cta(setup):
    """fp16 forward (mlp_fwd5) and fp16 chain (mlp_bwd_data5) keep two tiles in flight per CTA: 40 000 points = 313 tiles over 148
    CTAs (some CTAs get 3 tiles, the others 2: both slots, odd counts, ragged last tile) against the three-term path on the GPU."""
    cn = setup["cn"]
    packed, P = setup["packed"], setup["P"]
    gen = torch.Generator().manual_seed(11)
    n, S = 625, 64
    pts = (torch.randn(n, S, 3, generator=gen) * 1.5).to(DEV)
    vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1).to(DEV)
    d_raw = (torch.randn(n * S, 4, generator=gen) * 1e-6).to(DEV)
    raw3, acts3 = cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=3, fwd_terms=3)
    g3 = cn.ops.fused_mlp_backward(packed, P, acts3, d_raw, n * S, terms=(3, 3))
    del acts3
    for ft in (3, 1):
        raw1, acts1 = cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=1, fwd_terms=ft)
        assert rel_err(raw1, raw3) < (1e-3 if ft == 1 else 1e-7)
        g1 = cn.ops.fused_mlp_backward(packed, P, acts1, d_raw, n * S, terms=(1, 1))
        a = torch.cat([g1[k].double().reshape(-1) for k in sorted(g1)])
        b = torch.cat([g3[k].double().reshape(-1) for k in sorted(g3)])
        cos = float((a * b).sum() / (a.norm() * b.norm()))
        print(f"fwd_terms={ft}: fp16 chain + dW vs three-term, 40 000 points: 1 - cos = {1 - cos:.2e}")
        assert cos > (1.0 - 2e-3 if ft == 1 else 1.0 - 1e-5), (ft, cos)
        again = cn.ops.fused_mlp_backward(packed, P, acts1, d_raw, n * S, terms=(1, 1))
        assert all(torch.equal(again[k], g1[k]) for k in g1)                    # deterministic
        del acts1
