"""The drop-in boundary against the REAL reference script: import the unmodified
``run_nerf.py`` (I/O-only modules stubbed), patch its module globals, and let ITS OWN ``create_nerf`` build the networks,
the encoders, the query closure and the render kwargs from the product's classes; a checkpoint written in the reference
format is picked up by the script's own reload code.  CPU only (no kernel runs); skipped where the reference tree is absent
(the GPU box)."""
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/nerf-pytorch-master"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "run_nerf.py")), reason="reference tree not present")


def _args(basedir):
    return types.SimpleNamespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64, netdepth=8, netwidth=256,
        netdepth_fine=8, netwidth_fine=256, netchunk=1024 * 64, lrate=5e-4, basedir=str(basedir), expname="exp", ft_path=None,
        no_reload=False, perturb=1.0, white_bkgd=True, raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False,
        stable_init=False)


def test_reference_create_nerf_builds_product_objects(tmp_path):
    import consistentnerf_b200 as cn
    from consistentnerf_b200 import dropin, formats
    dropin.install_io_stubs()
    sys.path.insert(0, REF)
    try:
        import run_nerf as m                                   # the unmodified reference script
    finally:
        sys.path.remove(REF)
    patched = dropin.patch(m)
    for name in ("render", "render_rays", "batchify_rays", "raw2outputs", "run_network", "NeRF", "get_embedder", "sample_pdf"):
        assert name in patched
    assert m.NeRF is cn.NeRF and m.render_rays is not cn.render_rays        # run_nerf.py gets the 4-tuple flavour (no depth_map)
    os.makedirs(tmp_path / "exp")
    args = _args(tmp_path)
    render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer = m.create_nerf(args)       # reference code, our classes
    assert start == 0
    coarse, fine = render_kwargs_train["network_fn"], render_kwargs_train["network_fine"]
    assert isinstance(coarse, cn.NeRF) and isinstance(fine, cn.NeRF) and coarse.spec.is_canonical and fine.spec.is_canonical
    assert render_kwargs_train["N_importance"] == 128 and render_kwargs_train["ndc"] is False
    assert render_kwargs_test["perturb"] is False and render_kwargs_test["raw_noise_std"] == 0.0
    assert len(grad_vars) == len(list(coarse.parameters())) + len(list(fine.parameters()))
    assert callable(render_kwargs_train["network_query_fn"])
    # a reference-format checkpoint is found and loaded by the script's own reload code (NP/run_nerf.py:216-238)
    with torch.no_grad():
        coarse.pts_linears[0].bias.fill_(0.125)
    formats.save_checkpoint(str(tmp_path / "exp" / "000500.tar"), 500, coarse, fine, optimizer)
    kw2, _, start2, _, _ = m.create_nerf(args)
    assert start2 == 500
    assert torch.equal(kw2["network_fn"].pts_linears[0].bias, torch.full((256,), 0.125))



# run_nerf_view.py itself cannot be imported without a GPU (it calls torch.cuda.current_device() at import time,
# NP/run_nerf_view.py:39-40) and the GPU box has no reference tree, so the view flavour of the boundary is covered by
# test_distributed_cpu.test_dropin_patches_reference_globals and by the GPU tests that call the 5-tuple render().
