"""Host enqueue time of one training step vs its GPU time (is the step host bound?) + a cProfile of the enqueue path."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, consistentnerf_b200 as cn
from consistentnerf_b200.distributed import FlatGrads
dev = torch.device("cuda", 0)
coarse, fine = bench.make_nets(dev)
embed_fn, _ = cn.get_embedder(10, 0); embeddirs_fn, _ = cn.get_embedder(4, 0)
def query(i, v, f): return cn.run_network(i, v, f, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
kw = dict(network_query_fn=query, perturb=1.0, N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False, near=2.0, far=6.0)
hot = [p for net in (coarse, fine) for n_, p in net.named_parameters() if n_ in net.spec.param_names()]
flat = FlatGrads(hot); opt = torch.optim.Adam(hot, lr=5e-4, fused=True)
batch = tuple(x.to(dev) for x in bench.make_batch(4096, 0))
def step():
    o, d, tgt, prior, mask = batch
    flat.zero_()
    rgb, disp, acc, depth, ex = cn.render(1, 4096, None, chunk=32768, rays=(o, d), retraw=True, **kw)
    loss = (cn.masked_img_loss(rgb, tgt, mask, 0.2) + cn.masked_img_loss(ex["rgb0"], tgt, mask, 0.2) + cn.masked_depth_loss(depth, prior, mask, 6.0, 0.2, include_unmasked=True) + cn.masked_depth_loss(ex["depth0"], prior, mask, 6.0, 0.2, include_unmasked=True))
    loss.backward(); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
for trial in range(3):
    t0 = time.perf_counter()
    for _ in range(20): step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0)/20:.2f} ms/step, total {1e3*(t2-t0)/20:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
