"""Shared helpers of the parity tests: oracle parameters -> package modules, error metrics."""
import numpy as np
import torch

from oracle import nerf_oracle as O

ARCH = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=(4,), use_viewdirs=True)
SMALL = dict(D=4, W=64, input_ch=63, input_ch_views=0, output_ch=5, skips=(2,), use_viewdirs=False)


def module_from_params(params, arch, device="cuda"):
    """consistentnerf_b200.NeRF holding exactly the oracle's parameter dictionary."""
    import consistentnerf_b200 as cn
    net = cn.NeRF(D=arch["D"], W=arch["W"], input_ch=arch["input_ch"], input_ch_views=arch["input_ch_views"],
                  output_ch=arch["output_ch"], skips=list(arch["skips"]), use_viewdirs=arch["use_viewdirs"])
    sd = net.state_dict()
    for k, v in params.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        sd[k] = v.to(torch.float32)
    net.load_state_dict(sd)
    return net.to(device)


def workload_rays(n, seed=0, dtype=torch.float32):
    """Workload A of SURVEY.md section 8(d): o=(0,0,4)+0.1 N, d=normalize((0,0,-1)+0.2 N), near 2, far 6."""
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, 4.0]) + 0.1 * torch.randn(n, 3, generator=g)
    d = torch.tensor([0.0, 0.0, -1.0]) + 0.2 * torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return o.to(dtype), d.to(dtype)


def rel_err(a, b):
    """max |a-b| / max(|b|): the scale-relative error the 1e-4 bar is stated on."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rtol, atol, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True, err_msg=what)
