"""BASELINE config 5 on one GPU: time to render one 800 x 800 novel view (640 000 rays, 64 + 128 samples, chunk 32768) through
render(H, W, K, c2w=...), and through pipeline.render_path for 4 views (double-buffered D2H)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, consistentnerf_b200 as cn
dev = torch.device("cuda", 0)
coarse, fine = bench.make_nets(dev)
embed_fn, _ = cn.get_embedder(10, 0); embeddirs_fn, _ = cn.get_embedder(4, 0)
def query(i, v, f): return cn.run_network(i, v, f, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
kw = dict(network_query_fn=query, perturb=0.0, N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse, use_viewdirs=True,
          white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False, near=2.0, far=6.0)
H = W = 800
K = [[1111.1, 0.0, 400.0], [0.0, 1111.1, 400.0], [0.0, 0.0, 1.0]]
c2w = torch.tensor([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 4.0]], device=dev)
with torch.no_grad():
    for chunk in (32768, 131072):
        for _ in range(2): cn.render(H, W, K, chunk=chunk, c2w=c2w, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): cn.render(H, W, K, chunk=chunk, c2w=c2w, **kw)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"render 800x800, chunk {chunk:6d}: {dt * 1e3:7.1f} ms per image = {H * W / dt / 1e6:.3f} M rays/s")
    poses = torch.stack([c2w] * 4)
    cn.render_path(poses, (H, W, 1111.1), K, 32768, kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rgbs, disps, accs = cn.render_path(poses, (H, W, 1111.1), K, 32768, kw)
    dt = (time.perf_counter() - t0) / 4
    print(f"render_path, 4 views -> numpy on the host: {dt * 1e3:7.1f} ms per image = {H * W / dt / 1e6:.3f} M rays/s")
