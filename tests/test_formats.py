"""On-disk formats of the callers: reference-format checkpoints, PFM depth maps, metrics.txt.
CPU only.  Where the reference tree is present (the build container) its own ``read_pfm`` is executed on the files written
here; on the GPU box that cross-check is skipped and the round trips remain."""
import ast
import os
import re

import numpy as np
import pytest
import torch

from util import ARCH

REF_LOADER = "/root/reference/nerf-pytorch-master/load_blender.py"


def _net(cn):
    return cn.NeRF(D=ARCH["D"], W=ARCH["W"], input_ch=ARCH["input_ch"], input_ch_views=ARCH["input_ch_views"],
                   output_ch=ARCH["output_ch"], skips=list(ARCH["skips"]), use_viewdirs=True)


def _reference_read_pfm():
    """The unmodified read_pfm of the reference, lifted out of its module (whose other imports are not installed here)."""
    src = open(REF_LOADER).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "read_pfm")
    scope = {"np": np, "re": re}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF_LOADER, "exec"), scope)
    return scope["read_pfm"]


@pytest.mark.parametrize("shape", [(5, 7), (4, 6, 3), (3, 3, 1)])
def test_pfm_round_trip_and_reference_reader(tmp_path, shape):
    from consistentnerf_b200 import formats
    img = np.random.RandomState(0).rand(*shape).astype(np.float32) * 7.0
    path = str(tmp_path / "depth.pfm")
    formats.write_pfm(path, img, scale=1.0)
    back, scale = formats.read_pfm(path)
    assert scale == 1.0 and np.array_equal(back, img.reshape(back.shape))
    if os.path.exists(REF_LOADER):
        ref, ref_scale = _reference_read_pfm()(path)
        assert ref_scale == 1.0 and np.array_equal(ref, back)


def test_checkpoint_reference_format_and_reload_semantics(tmp_path):
    import consistentnerf_b200 as cn
    from consistentnerf_b200 import formats
    torch.manual_seed(0)
    coarse, fine = _net(cn), _net(cn)
    opt = torch.optim.Adam(list(coarse.parameters()) + list(fine.parameters()), lr=5e-4)
    path = formats.save_checkpoint(str(tmp_path / "exp" / "{:06d}.tar".format(2500)), 2500, coarse, fine, opt)
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"global_step", "network_fn_state_dict", "network_fine_state_dict", "optimizer_state_dict"}
    assert formats.latest_checkpoint(str(tmp_path / "exp")) == path
    c2, f2 = _net(cn), _net(cn)
    start = formats.load_checkpoint(path, c2, f2)                       # create_nerf semantics (NP/run_nerf_view.py:350-363)
    assert start == 2500
    for k, v in coarse.state_dict().items():
        if k in ("temp_rgb", "temp_depth", "depth_scale"):
            assert float(c2.state_dict()[k]) == pytest.approx(0.1)      # reset on reload, as the reference does
        else:
            assert torch.equal(c2.state_dict()[k], v)
    c3 = _net(cn)
    formats.load_checkpoint(path, c3, None, reference_reload_semantics=False)
    assert float(c3.state_dict()["temp_rgb"]) == pytest.approx(-0.7)


def test_metrics_file(tmp_path):
    from consistentnerf_b200 import formats
    path = str(tmp_path / "metrics.txt")
    formats.write_metrics(path, 21.5, torch.tensor(0.83), 0.125)
    text = open(path).read()
    assert text.startswith("PSNR: 21.5\nSSIM: ") and text.endswith("LPIPS: 0.125") and text.count("\n") == 2
    m = formats.read_metrics(path)
    assert m["PSNR"] == 21.5 and abs(m["SSIM"] - 0.83) < 1e-6 and m["LPIPS"] == 0.125
