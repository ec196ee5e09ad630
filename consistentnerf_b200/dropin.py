"""Swap the hot path into the UNMODIFIED reference scripts.

The reference scripts resolve ``render``, ``render_rays``, ``NeRF`` ... by module-global lookup
(SURVEY.md section 8b), so replacing those globals is all a drop-in needs:

    import run_nerf_view as m                      # unmodified reference script
    from consistentnerf_b200 import dropin
    dropin.patch(m)                                # renderer, model, encoders, cross-view geometry
    torch.set_default_tensor_type('torch.cuda.FloatTensor')   # what the script's __main__ does (:2306)
    m.train()

``python -m consistentnerf_b200.dropin run_nerf_view --config ...`` does exactly that, after putting
stand-ins for the I/O-only modules this image lacks (imageio, configargparse, lpips ...) into sys.modules
when they are not importable.
"""
from __future__ import annotations

import importlib
import sys
import types

import torch

from . import consistency, nerf
from .render import make_api        # (the package attribute `render` is the function, not the submodule)

# names replaced in a reference script module, grouped by origin
_RENDER_NAMES = ["batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays"]
_MODEL_NAMES = ["NeRF", "Embedder", "get_embedder", "sample_pdf", "get_rays", "ndc_rays"]
_VIEW_NAMES = ["get_rays_ref", "get_ref_rays", "get_test_label"]


def wants_depth(module) -> bool:
    """The view scripts return depth_map from render(); run_nerf.py does not (NP/run_nerf.py:134 vs
    NP/run_nerf_view.py:246)."""
    return hasattr(module, "get_ref_rays") or "view" in getattr(module, "__name__", "")


def patch(module, with_depth=None):
    """Replace the hot-path globals of an imported reference script; returns the list of names patched."""
    if with_depth is None:
        with_depth = wants_depth(module)
    api = make_api(with_depth)
    done = []
    for name in _RENDER_NAMES:
        if hasattr(module, name):
            setattr(module, name, getattr(api, name))
            done.append(name)
    for name in _MODEL_NAMES:
        if hasattr(module, name):
            setattr(module, name, getattr(nerf, name))
            done.append(name)
    for name in _VIEW_NAMES:
        if hasattr(module, name):
            setattr(module, name, getattr(consistency, name))
            done.append(name)
    return done


def install_io_stubs():
    """Empty stand-ins for modules the reference imports only for I/O / logging when they are absent."""
    for name in ["imageio", "ipdb", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "configargparse",
                 "tensorboardX", "pytorch_msssim", "lpips"]:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "__path__"):
        mpl.__path__ = []
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        mpl.cm = sys.modules["matplotlib.cm"]
    lp = sys.modules["lpips"]
    if not hasattr(lp, "LPIPS"):
        class _NoLPIPS:
            def __init__(self, *a, **k):
                pass

            def to(self, *a, **k):
                return self

            def __call__(self, *a, **k):
                raise RuntimeError("lpips is not installed in this image")
        lp.LPIPS = _NoLPIPS
    tb = sys.modules["tensorboardX"]
    if not hasattr(tb, "SummaryWriter"):
        class _NullWriter:
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, _):
                return lambda *a, **k: None
        tb.SummaryWriter = _NullWriter
    ms = sys.modules["pytorch_msssim"]
    for n in ("ssim", "ms_ssim"):
        if not hasattr(ms, n):
            setattr(ms, n, lambda *a, **k: (_ for _ in ()).throw(RuntimeError("pytorch_msssim is not installed")))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: python -m consistentnerf_b200.dropin <reference_script_module> [script args...]")
        return 2
    script, rest = argv[0], argv[1:]
    install_io_stubs()
    module = importlib.import_module(script)
    patched = patch(module)
    print(f"[consistentnerf_b200] patched {script}: {', '.join(patched)}")
    torch.set_default_tensor_type("torch.cuda.FloatTensor")
    sys.argv = [script + ".py"] + rest
    return module.train()


if __name__ == "__main__":
    sys.exit(main() or 0)
