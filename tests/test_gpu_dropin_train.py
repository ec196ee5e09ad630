"""The UNMODIFIED reference scripts, driven through the drop-in on the GPU (missing #1 / weak #1):

* ``run_nerf.py`` train() (Blender loader -> render -> img2mse -> backward -> Adam, NP/run_nerf.py:537-877) and
  ``run_nerf_view.py`` train() (DTU loader, hard-mask loop through get_ref_rays, masked rgb + depth losses, clip_grad_value_,
  NP/run_nerf_view.py:811-2300) run N iterations with the hot path patched in, under the scripts' own
  ``torch.set_default_tensor_type('torch.cuda.FloatTensor')`` regime, on synthetic scenes written in the reference's formats.
* the product's public functions give the same results under that default tensor type as under the normal one.

The scripts come from oracle/_ref (oracle/build_ref.py; shipped to the GPU box); each run is a subprocess (oracle/twin.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _twin():
    sys.path.insert(0, ROOT)
    from oracle import build_ref, twin
    if not build_ref.available():
        pytest.skip("oracle/_ref not populated (python oracle/build_ref.py where /root/reference exists)")
    return twin


def test_run_nerf_train_through_dropin(tmp_path):
    twin = _twin()
    root = str(tmp_path / "blender")
    twin.make_scene("blender", root, res=96)
    res = twin.run_arm_subprocess("repo", "blender", root, iters=40, eval_views=2, extra_args=twin.FULL[:-1] + ["1024"], timeout=900)
    assert "error" not in res, res
    assert {"render", "render_rays", "batchify_rays", "raw2outputs", "run_network", "NeRF", "get_embedder", "sample_pdf", "get_rays"} <= set(res["patched"])
    assert res["checkpoints"] == ["000040.tar"]                       # written by the script's own torch.save (NP/run_nerf.py:797-805)
    assert res["psnr"] > 8.0 and res["train_ms_per_iter"] > 0          # finite numbers: 40 iterations only prove the plumbing
    log = open(os.path.join(root, "log_blender_repo.txt")).read()
    assert "[TRAIN] Iter:" in log and "nan" not in log.lower().split("[train] iter:")[-1]


def test_run_nerf_view_dtu_hardmask_depth_loss_through_dropin(tmp_path):
    """BASELINE config 3 recipe: --hardmask --with_depth_loss --no_batching on DTU-format data; the 327 680-pixel hard-mask loop
    runs through the patched get_ref_rays (64 chunks of 5120 per pair, NP/run_nerf_view.py:1014-1041)."""
    twin = _twin()
    root = str(tmp_path / "dtu")
    twin.make_scene("dtu", root)
    res = twin.run_arm_subprocess("repo", "dtu", root, iters=25, eval_views=1, eval_res_div=4, extra_args=twin.FULL[:-1] + ["1024"], timeout=1500)
    assert "error" not in res, res
    assert {"render", "render_rays", "get_ref_rays", "get_test_label", "NeRF"} <= set(res["patched"])
    assert res["checkpoints"] == ["000025.tar"]
    masks = [f for f in os.listdir(os.path.join(root, "logs", "twin_dtu", "mask", "scan114", "3view")) if f.endswith(".jpg")]
    assert len(masks) == 49                                            # one mask image per view, written by the script
    assert res["psnr"] > 5.0


_REGIME = r"""
import sys, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch
import consistentnerf_b200 as cn
from consistentnerf_b200 import pipeline, consistency
import bench

dev = torch.device("cuda")
coarse, fine = bench.make_nets(dev)                    # built and drawn ONCE, before the regime switch: same weights, same inputs
o, d = bench.workload_rays(256, 0)
cpu_rand = lambda seed, *shape: torch.rand(*shape, generator=torch.Generator(device="cpu").manual_seed(seed), device="cpu").to(dev)
w_pdf, x_in, tgt, pw = cpu_rand(1, 7, 62), cpu_rand(2, 33, 90), cpu_rand(3, 256, 3), cpu_rand(4, 40, 3) * 2 - 1
imgs = np.random.RandomState(0).rand(2, 20, 24, 3).astype(np.float32)

def run():
    e, _ = cn.get_embedder(10, 0); ev, _ = cn.get_embedder(4, 0)
    q = lambda i, v, f: cn.run_network(i, v, f, embed_fn=e, embeddirs_fn=ev)
    kw = dict(network_query_fn=q, perturb=0.0, N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse, use_viewdirs=True,
              white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False, near=2.0, far=6.0)
    out = {}
    with torch.no_grad():
        r = cn.render(1, 256, None, chunk=100, rays=(o.to(dev), d.to(dev)), retraw=True, **kw)
    out["render_rays"] = [r[0], r[1], r[2], r[3], r[4]["rgb0"], r[4]["z_std"], r[4]["raw"]]
    K = np.array([[40.0, 0, 12], [0, 40.0, 10], [0, 0, 1]])
    c2w = torch.Tensor([[1, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, 4.0]])
    with torch.no_grad():
        r = cn.render(20, 24, K, chunk=4096, c2w=c2w, **kw)
    out["render_c2w"] = [r[0], r[3]]
    ro, rd = cn.get_rays(20, 24, K, c2w)
    out["get_rays"] = [ro, rd]
    rgbs, disps, accs = pipeline.render_path([c2w, c2w], [20, 24, 40.0], K, 4096, kw)
    out["render_path"] = [torch.from_numpy(rgbs), torch.from_numpy(accs)]
    bank = pipeline.RayBank(imgs, np.stack([np.eye(4)[:3], np.eye(4)[:3]]), K, 2.0, 6.0, depths=imgs[..., 0], masks=(imgs[..., 1] > 0.5).astype(np.float32))
    b = bank.sample(64, view=1)
    out["raybank"] = [b["rays"], b["target"], b["depth"], b["mask"]]
    bins = torch.linspace(2, 6, 63, device="cpu")[None].repeat(7, 1).to(dev)
    out["sample_pdf"] = [cn.sample_pdf(bins, w_pdf, 16, det=True)]
    with torch.no_grad():
        out["nerf_forward"] = [coarse(x_in)]
    m = (tgt[:, :1] > 0.4).float()
    rr = cn.render(1, 256, None, chunk=4096, rays=(o.to(dev), d.to(dev)), retraw=True, **kw)
    loss = consistency.masked_img_loss(rr[0], tgt, m, 0.2) + consistency.masked_depth_loss(rr[3], tgt[:, 0] * 4, m, 6.0) + cn.img2mse(rr[4]["rgb0"], tgt)
    for p in list(coarse.parameters()) + list(fine.parameters()):
        p.grad = None
    loss.backward()
    out["train"] = [loss.detach(), fine.pts_linears[3].weight.grad, coarse.alpha_linear.bias.grad]
    w2c = torch.eye(4)[None]; w2c[0, 2, 3] = 4.0
    img = torch.from_numpy(imgs[:1]).permute(0, 3, 1, 2).to(dev)
    g = consistency.get_test_label(w2c.to(dev), w2c.to(dev), torch.Tensor(K)[None].to(dev), pw[None, :, None, :], img)
    out["test_label"] = [g[0], g[1], g[2].float(), g[3]]
    return {k: [t.detach().float().cpu() for t in v] for k, v in out.items()}

a = run()
torch.set_default_tensor_type("torch.cuda.FloatTensor")          # what the reference scripts' __main__ does (NP/run_nerf_view.py:2306)
b = run()
bad = []
for k in a:
    for i, (x, y) in enumerate(zip(a[k], b[k])):
        if x.shape != y.shape or not torch.equal(torch.nan_to_num(x), torch.nan_to_num(y)):
            bad.append((k, i, float((x - y).abs().max()) if x.shape == y.shape else "shape"))
print("REGIME", json.dumps({"checked": sorted(a), "bad": bad}))
"""


def test_public_functions_under_default_cuda_tensor_type():
    """Every public entry point gives bit-identical results with and without the CUDA default tensor type."""
    res = subprocess.run([sys.executable, "-c", _REGIME % (ROOT, os.path.join(ROOT, "tests"))], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("REGIME")][-1]
    info = json.loads(line[len("REGIME "):])
    assert len(info["checked"]) >= 9 and info["bad"] == [], info
