"""K7b soft-weighted MSE (NP/run_nerf_view.py:50-58) without a GPU.

1. the oracle's restatement equals the reference's own lambdas (their source lines, evaluated from oracle/_ref/run_nerf_view.py --
   the module itself cannot be imported on a CPU: it moves an LPIPS network to CUDA at import, :40);
2. the element arithmetic the CUDA kernels are built from (csrc/soft_weight.h), compiled for the host against the prototypes of
   include/cnerf.h (tests/host/soft_mse_host.c), driven through the package's real autograd glue (ops.SoftMSEFn, consistency.img2mse_*)
   with only the library call redirected, gives the oracle's loss and gradients (pred, and temp for kind 0).
The CUDA kernels themselves are checked on the GPU (tests/test_gpu_zz_soft_mse.py)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from oracle import nerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SCRIPT = os.path.join(ROOT, "oracle", "_ref", "run_nerf_view.py")


def reference_lambdas():
    ns = {"torch": torch}
    for line in open(REF_SCRIPT):
        if re.match(r"^(img2mse_softmask|img2mse_depth_softmask|img2mse_softLpmask) = lambda", line):
            exec(line, ns)
    return ns


def cases(seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(300, 3, generator=g, dtype=torch.float64)
    y = torch.rand(300, 3, generator=g, dtype=torch.float64)
    return x, y


@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="oracle/_ref not populated (python oracle/build_ref.py)")
def test_oracle_equals_the_reference_lambdas():
    ns = reference_lambdas()
    assert {"img2mse_softmask", "img2mse_depth_softmask", "img2mse_softLpmask"} <= set(ns)
    x, y = cases()
    for name, kind, param in (("img2mse_softmask", 0, 0.4), ("img2mse_depth_softmask", 0, 1.3), ("img2mse_softLpmask", 1, 2.0),
                              ("img2mse_softLpmask", 1, 1.5)):
        grads = []
        for fn in (lambda a, b, p: ns[name](a, b, p), lambda a, b, p: O.soft_mse(a, b, p, kind)):
            a = x.clone().requires_grad_(True)
            p = torch.tensor(param, dtype=torch.float64, requires_grad=(kind == 0))
            loss = fn(a, y, p if kind == 0 else param)
            loss.backward()
            grads.append((loss.detach(), a.grad, p.grad))
        assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1]), name
        if kind == 0:
            assert torch.equal(grads[0][2], grads[1][2]), name


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("soft") / "soft_mse_host.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "consistentnerf_b200", "csrc"), os.path.join(ROOT, "tests", "host", "soft_mse_host.c"),
                    "-o", so, "-lm"], check=True)
    return ctypes.CDLL(so)


@pytest.fixture()
def redirected(host_lib, monkeypatch):
    """ops.call -> the host build for the two K7b names (any other name would be a bug here); CPU tensors allowed through."""
    from consistentnerf_b200 import _lib, ops
    calls = []

    def call(name, *args):
        assert name in ("cnerf_soft_mse_fwd", "cnerf_soft_mse_bwd"), name
        fn = getattr(host_lib, name)
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]
        calls.append(name)
        assert fn(*args) == 0

    monkeypatch.setattr(ops, "call", call)
    monkeypatch.setattr(ops, "ptr", lambda t: None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr()))
    monkeypatch.setattr(ops, "stream", lambda: None)
    monkeypatch.setattr(ops, "_need_cuda", lambda t, what: None)
    monkeypatch.setattr(ops, "_workspace", lambda device, n: torch.empty(n, dtype=torch.uint8))
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))     # (the glue insists on a device-resident temp)
    return calls


@pytest.mark.parametrize("kind,param", [(0, 0.4), (0, 1.3), (1, 2.0), (1, 1.5), (1, 1.0)])
def test_kernel_arithmetic_and_autograd_glue_against_the_oracle(redirected, kind, param):
    from consistentnerf_b200 import consistency
    x64, y64 = cases(3 + kind)
    a64 = x64.clone().requires_grad_(True)
    p64 = torch.tensor([param], dtype=torch.float64, requires_grad=(kind == 0))
    want = O.soft_mse(a64, y64, p64 if kind == 0 else param, kind)
    (3.0 * want).backward()

    a = x64.float().requires_grad_(True)
    p = torch.tensor([param], requires_grad=(kind == 0))
    fn = consistency.img2mse_softmask if kind == 0 else consistency.img2mse_softLpmask
    got = fn(a, y64.float(), p if kind == 0 else param)
    (3.0 * got).backward()
    assert redirected == ["cnerf_soft_mse_fwd", "cnerf_soft_mse_bwd"]
    assert got.shape == () and a.grad.shape == a.shape
    assert abs(float(got) - float(want)) <= 2e-6 * abs(float(want))
    assert float((a.grad.double() - a64.grad).abs().max()) <= 5e-6 * float(a64.grad.abs().max())
    if kind == 0:
        assert p.grad.shape == p.shape
        assert abs(float(p.grad) - float(p64.grad)) <= 1e-4 * abs(float(p64.grad)) + 1e-9
        # python-float temp: same loss, no parameter gradient
        got2 = consistency.img2mse_depth_softmask(x64.float(), y64.float(), param)
        assert abs(float(got2) - float(want)) <= 2e-6 * abs(float(want))


def test_depth_form_divides_inside(redirected):
    from consistentnerf_b200 import consistency
    x64, y64 = cases(9)
    far = 5.5
    d64 = (x64[:, 0] * 6).clone().requires_grad_(True)
    want = O.soft_mse(d64 / far, y64[:, 0] * 6 / far, 2.0, 1)
    want.backward()
    d = (x64[:, 0] * 6).float().requires_grad_(True)
    got = consistency.soft_depth_loss(d, (y64[:, 0] * 6).float(), far, 2.0, kind=1)
    got.backward()
    assert abs(float(got) - float(want)) <= 2e-6 * abs(float(want))
    assert float((d.grad.double() - d64.grad).abs().max()) <= 5e-6 * float(d64.grad.abs().max())


def test_empty_input_is_nan_like_the_reference(redirected):
    from consistentnerf_b200 import consistency
    e = torch.zeros(0, 3)
    assert torch.isnan(consistency.img2mse_softLpmask(e, e, 2.0))
    assert torch.isnan(O.soft_mse(e.double(), e.double(), 2.0, 1))


def test_dropin_patches_the_loss_names():
    import types
    from consistentnerf_b200 import consistency, dropin
    m = types.SimpleNamespace(__name__="run_nerf_view", img2mse_softmask=None, img2mse_depth_softmask=None, img2mse_softLpmask=None,
                              get_ref_rays=None)
    done = dropin.patch(m)
    assert {"img2mse_softmask", "img2mse_depth_softmask", "img2mse_softLpmask", "get_ref_rays"} <= set(done)
    assert m.img2mse_softLpmask is consistency.img2mse_softLpmask


GOLD = (("exp_rgb", 0), ("exp_depth", 0), ("lp2", 1), ("lp15", 1))


def test_oracle_against_the_reference_golden():
    from conftest import load_golden, t
    g = load_golden("soft_mse")
    for tag, kind in GOLD:
        a = t(g["x"]).double().requires_grad_(True)
        p = torch.tensor([float(g[tag + "_param"])], dtype=torch.float64, requires_grad=(kind == 0))
        loss = O.soft_mse(a, t(g["y"]).double(), p if kind == 0 else float(p), kind)
        loss.backward()
        assert abs(float(loss.detach()) - float(g[tag + "_loss"][0])) <= 2e-6 * float(g[tag + "_loss"][0]), tag
        assert float((a.grad - t(g[tag + "_dx"]).double()).abs().max()) <= 1e-5 * float(a.grad.abs().max()), tag
        if kind == 0:
            assert abs(float(p.grad) - float(g[tag + "_dparam"][0])) <= 1e-4 * abs(float(p.grad)), tag


def test_host_build_against_the_reference_golden(redirected):
    from conftest import load_golden, t
    from consistentnerf_b200 import consistency
    g = load_golden("soft_mse")
    for tag, kind in GOLD:
        a = t(g["x"]).requires_grad_(True)
        param = float(g[tag + "_param"])
        p = torch.tensor([param], requires_grad=True) if kind == 0 else param
        loss = (consistency.img2mse_softmask if kind == 0 else consistency.img2mse_softLpmask)(a, t(g["y"]), p)
        loss.backward()
        assert abs(float(loss.detach()) - float(g[tag + "_loss"][0])) <= 2e-6 * float(g[tag + "_loss"][0]), tag
        assert float((a.grad - t(g[tag + "_dx"])).abs().max()) <= 1e-5 * float(a.grad.abs().max()), tag
        if kind == 0:
            assert abs(float(p.grad) - float(g[tag + "_dparam"][0])) <= 1e-4 * abs(float(p.grad)), tag
