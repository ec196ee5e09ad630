"""Diagnostic: cnerf_sample_pdf debug outputs (cdf, below) against the golden vectors and the host oracle (ISA dependence of torch.sum)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden, t
from oracle import nerf_oracle as O
from util import *
import consistentnerf_b200 as cn
DEV = "cuda"
g = load_golden("sample_pdf")
for mode in ("det", "rnd"):
    s, dbg = cn.ops.sample_pdf(t(g["bins"], device=DEV), t(g["weights"], device=DEV), t(g["u_" + mode], device=DEV), 128, debug=True)
    cdf = dbg["cdf"].cpu(); gc = t(g["cdf_" + mode])
    print(mode, "cdf max abs diff", float((cdf - gc).abs().max()), "n diff", int((cdf != gc).sum()), "of", cdf.numel())
    inds = t(g["inds_" + mode]); below = torch.clamp(inds - 1, min=0)
    nb = int((dbg["below"].cpu().long() != below).sum())
    print(mode, "below mismatches", nb, "samples max diff", float((s.cpu() - t(g["samples_" + mode])).abs().max()))
    # host oracle vs golden (ISA dependence)
    so, do = O.sample_pdf(t(g["bins"]), t(g["weights"]), t(g["u_" + mode]), return_debug=True)
    print(mode, "host-oracle cdf == golden:", bool(torch.equal(do["cdf"], gc)), torch.backends.cpu.get_cpu_capability())

# gradient diag
from test_gpu_render import _kwargs
n = 96
o, d = workload_rays(n, seed=3)
pc = O.make_params(2, sigma_bias=0.5, **ARCH); pf = O.make_params(3, sigma_bias=0.5, **ARCH)
coarse, fine = module_from_params(pc, ARCH), module_from_params(pf, ARCH)
tgt = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
for noise_std in (0.0, 1.0):
    for net in (coarse, fine):
        net.zero_grad()
    kw = _kwargs(cn, coarse, fine, perturb=1.0, raw_noise_std=noise_std, pytest=True)
    rgb, disp, acc, depth, ex = cn.render(1, n, None, chunk=4096, rays=(o.to(DEV), d.to(DEV)), retraw=True, **kw)
    loss = cn.img2mse(rgb, tgt.to(DEV)) + cn.img2mse(ex["rgb0"], tgt.to(DEV))
    loss.backward()
    np.random.seed(0); t_rand = torch.tensor(np.random.rand(n, 64))
    np.random.seed(0); noise_c = torch.tensor(np.random.rand(n, 64)) * noise_std
    np.random.seed(0); u = torch.tensor(np.random.rand(n, 128))
    np.random.seed(0); noise_f = torch.tensor(np.random.rand(n, 192)) * noise_std
    for dt in (torch.float64, torch.float32):
        c64 = {k: v.to(dt).requires_grad_(True) for k, v in pc.items()}
        f64 = {k: v.to(dt).requires_grad_(True) for k, v in pf.items()}
        rays64 = O.pack_rays(o.to(dt), d.to(dt), 2.0, 6.0, True)
        ref = O.render_rays(rays64, c64, f64, ARCH, n_samples=64, n_importance=128, white_bkgd=True, t_rand=t_rand.float().to(dt), u=u.float().to(dt),
                            noise_coarse=noise_c.float().to(dt) if noise_std else None, noise_fine=noise_f.float().to(dt) if noise_std else None)
        ref_loss = ((ref["rgb_map"] - tgt.to(dt)) ** 2).mean() + ((ref["rgb0"] - tgt.to(dt)) ** 2).mean()
        ref_loss.backward()
        if dt == torch.float64:
            base = (c64, f64)
            print("noise", noise_std, "loss", float(loss.detach()), float(ref_loss.detach()))
            for tag, net, p64 in (("coarse", coarse, c64), ("fine", fine, f64)):
                for name, prm in net.named_parameters():
                    if p64[name].grad is not None and "weight" in name:
                        print(f"  {tag:6s} {name:24s} cand {rel_err(prm.grad, p64[name].grad):.2e}  max|g| {float(p64[name].grad.abs().max()):.2e}")
        else:
            for tag, p32, p64 in (("coarse", c64, base[0]), ("fine", f64, base[1])):
                for name in ("pts_linears.0.weight", "pts_linears.7.weight", "rgb_linear.weight"):
                    print(f"  fp32-oracle {tag} {name} {rel_err(p32[name].grad, p64[name].grad):.2e}")
