"""BASELINE config 4 (LLFF fern, 6 training views, NDC): the UNMODIFIED run_nerf_view.py train() through the drop-in on an LLFF-format
scene written by this package (oracle/twin.make_llff: poses_bounds.npy in LLFF axis order, images_8/, prior depths, pairs) -- the
reference's own load_llff_data (rescale, recentre, spiral path), the 190 512-pixel hard-mask loop (38 chunks of 5120 per pair, 30
pairs), training and held-out rendering in NDC (near 0 / far 1, ndc_rays inside the patched render), raw_noise_std = 1.

The same scene and harness are exercised on the CPU with the reference arm (tests/test_dataset_formats.py::
test_llff_reference_arm_trains_on_the_cpu).  This GPU arm was written after the round's GPU budget was spent: its first run on a
B200 is the round-end driver's, hence the non-strict xfail -- an XPASS is the expected outcome, a failure informs without gating."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="first GPU run of this arm happens at round end (GPU budget spent); CPU-validated on the reference arm")
def test_run_nerf_view_llff_ndc_hardmask_through_dropin(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import build_ref, twin
    if not build_ref.available():
        pytest.skip("oracle/_ref not populated (python oracle/build_ref.py where /root/reference exists)")
    root = str(tmp_path / "llff")
    twin.make_scene("llff", root)
    res = twin.run_arm_subprocess("repo", "llff", root, iters=25, eval_views=1, eval_res_div=2, extra_args=twin.FULL[:-1] + ["1024"], timeout=1500)
    assert "error" not in res, res
    assert {"render", "render_rays", "get_rays", "ndc_rays", "get_ref_rays", "NeRF"} <= set(res["patched"])
    assert res["checkpoints"] == ["000025.tar"]
    masks = [f for f in os.listdir(os.path.join(root, "logs", "twin_fern", "mask", "fern", "6view")) if f.endswith(".jpg")]
    assert len(masks) == 20
    log = open(os.path.join(root, "log_llff_repo.txt")).read()
    assert "NEAR FAR 0.0 1.0" in log and "[TRAIN] Iter:" in log
    assert res["psnr"] > 5.0
