"""Small end-to-end exercise of every tcgen05 / TMA kernel for compute-sanitizer (memcheck, racecheck, synccheck):

    compute-sanitizer --tool memcheck  python scripts/sanitize.py
    compute-sanitizer --tool racecheck python scripts/sanitize.py

300 points (two full 128-point tiles + a ragged one) through: three-term forward (inference / training), fp16 two-tile forward
(inference / training), data-gradient chain (three-term and fp16), weight-gradient GEMMs (three-term and fp16 with wide stages),
narrow heads, plus the per-ray kernels via one tiny render() with gradients."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import consistentnerf_b200 as cn

dev = torch.device("cuda")
coarse, fine = bench.make_nets(dev)
packed = coarse.packed_weights()
P = dict(zip(coarse.spec.param_names(), [p.detach() for p in coarse.hot_params()]))
packed.refresh(P)
n, S = 5, 60
g = torch.Generator().manual_seed(0)
pts = (torch.randn(n, S, 3, generator=g) * 1.5).to(dev)
vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
d_raw = (torch.randn(n * S, 4, generator=g) * 1e-6).to(dev)
for ft in (3, 1):
    cn.ops.fused_mlp_forward(packed, pts, vd, fwd_terms=ft)
for ft, terms in ((3, (3, 3)), (3, (3, 1)), (3, (1, 1)), (1, (1, 1))):
    raw, acts = cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=terms[1], fwd_terms=ft)
    cn.ops.fused_mlp_backward(packed, P, acts, d_raw, n * S, terms=terms)
torch.cuda.synchronize()
e, _ = cn.get_embedder(10, 0)
ev, _ = cn.get_embedder(4, 0)
q = lambda i, v, f: cn.run_network(i, v, f, embed_fn=e, embeddirs_fn=ev)
kw = dict(network_query_fn=q, perturb=1.0, N_importance=16, network_fine=fine, N_samples=8, network_fn=coarse, use_viewdirs=True,
          white_bkgd=True, raw_noise_std=1.0, ndc=False, lindisp=False, near=2.0, far=6.0)
o, d = bench.workload_rays(24, 0)
rgb, disp, acc, depth, ex = cn.render(1, 24, None, chunk=16, rays=(o.to(dev), d.to(dev)), retraw=True, **kw)
tgt = torch.rand(24, 3, device=dev)
m = (tgt[:, :1] > 0.5).float()
loss = cn.masked_img_loss(rgb, tgt, m, 0.2) + cn.masked_depth_loss(depth, tgt[:, 0] * 4, m, 6.0) + cn.img2mse(ex["rgb0"], tgt)
loss.backward()
torch.cuda.synchronize()
print("sanitize.py: done, loss", float(loss))
