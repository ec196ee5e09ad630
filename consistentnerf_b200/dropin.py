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

import torch

from . import consistency, nerf
from .render import make_api        # (the package attribute `render` is the function, not the submodule)

# names replaced in a reference script module, grouped by origin
_RENDER_NAMES = ["batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays"]
_MODEL_NAMES = ["NeRF", "Embedder", "get_embedder", "sample_pdf", "get_rays", "ndc_rays"]
_VIEW_NAMES = ["get_rays_ref", "get_ref_rays", "get_test_label"]
_LOSS_NAMES = ["img2mse_softmask", "img2mse_depth_softmask", "img2mse_softLpmask"]      # --softLpmask (NP/run_nerf_view.py:1663)


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
    for name in _VIEW_NAMES + _LOSS_NAMES:
        if hasattr(module, name):
            setattr(module, name, getattr(consistency, name))
            done.append(name)
    return done


def install_io_stubs():
    """Make the I/O-only modules the reference imports resolvable when they are absent (imageio, configargparse, matplotlib,
    tensorboardX, ipdb, pytorch_msssim, lpips): functional stand-ins from ``consistentnerf_b200/shims`` are appended to
    ``sys.path``, so an installed original always wins."""
    from . import shims
    return shims.install()


def bound_iterations(module, max_iters: int):
    """Stop the script's training loop after ``max_iters`` steps: run_nerf.py hard-codes 200 001 iterations (NP/run_nerf.py:704)
    and iterates with the module-global ``trange``."""
    def bounded(a, b=None, *args, **kwargs):
        lo, hi = (0, a) if b is None else (a, b)
        return range(lo, min(hi, lo + int(max_iters)))
    module.trange = bounded


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: python -m consistentnerf_b200.dropin <reference_script_module> [--cnerf_max_iters N] [script args...]")
        return 2
    script, rest = argv[0], argv[1:]
    max_iters = None
    if "--cnerf_max_iters" in rest:                   # launcher option, not passed on to the script
        i = rest.index("--cnerf_max_iters")
        max_iters = int(rest[i + 1])
        del rest[i:i + 2]
    install_io_stubs()
    module = importlib.import_module(script)
    patched = patch(module)
    print(f"[consistentnerf_b200] patched {script}: {', '.join(patched)}")
    if max_iters is not None:
        bound_iterations(module, max_iters)
    torch.set_default_tensor_type("torch.cuda.FloatTensor")
    sys.argv = [script + ".py"] + rest
    return module.train()


if __name__ == "__main__":
    sys.exit(main() or 0)
