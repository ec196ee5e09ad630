"""Dataset directories of SURVEY.md section 8f item 4: the writers of consistentnerf_b200/formats.py produce what the
REFERENCE's own loaders read (NP/load_blender.py, NP/load_dtu.py, NP/load_llff.py -- imported unmodified from oracle/_ref or
/root/reference when present), and the native loaders return the same arrays.  Also the shims that stand in for the I/O-only
modules (imageio, configargparse ...).  CPU only."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, "oracle", "_ref"), "/root/reference/nerf-pytorch-master"]
REF = next((d for d in REF_DIRS if os.path.exists(os.path.join(d, "load_blender.py"))), None)
needs_ref = pytest.mark.skipif(REF is None, reason="reference loaders not available (run oracle/build_ref.py)")


def _import_ref(name):
    from consistentnerf_b200 import shims
    shims.install()
    sys.path.insert(0, REF)
    try:
        return __import__(name)
    finally:
        sys.path.remove(REF)


def _scene(n, h, w, seed=0):
    rng = np.random.RandomState(seed)
    imgs = (rng.rand(n, h, w, 4) * 255).astype(np.uint8)
    poses = np.tile(np.eye(4), (n, 1, 1))
    poses[:, :3, 3] = rng.randn(n, 3)
    return imgs, poses


def test_blender_roundtrip_native(tmp_path):
    from consistentnerf_b200 import formats
    splits = {k: _scene(n, 6, 8, i) for i, (k, n) in enumerate((("train", 3), ("val", 1), ("test", 2)))}
    formats.write_blender_scene(str(tmp_path), splits, camera_angle_x=0.7)
    imgs, poses, hwf, i_split = formats.load_blender_scene(str(tmp_path))
    assert imgs.shape == (6, 6, 8, 4) and [len(s) for s in i_split] == [3, 1, 2]
    assert np.array_equal((imgs[:3] * 255 + 0.5).astype(np.uint8), splits["train"][0])
    np.testing.assert_allclose(poses[3], splits["val"][1][0], atol=1e-6)
    assert hwf[:2] == [6, 8] and abs(hwf[2] - 0.5 * 8 / np.tan(0.35)) < 1e-9
    meta = json.load(open(tmp_path / "transforms_test.json"))
    assert set(meta) == {"camera_angle_x", "frames"} and set(meta["frames"][0]) >= {"file_path", "transform_matrix"}


@needs_ref
def test_blender_scene_is_read_by_the_reference_loader(tmp_path):
    from consistentnerf_b200 import formats
    lb = _import_ref("load_blender")
    splits = {k: _scene(n, 6, 8, i) for i, (k, n) in enumerate((("train", 3), ("val", 1), ("test", 2)))}
    formats.write_blender_scene(str(tmp_path), splits, camera_angle_x=0.7)
    imgs, poses, render_poses, hwf, i_split = lb.load_blender_data(str(tmp_path), half_res=False, testskip=1)
    mine = formats.load_blender_scene(str(tmp_path))
    assert np.array_equal(imgs, mine[0]) and np.array_equal(poses, mine[1]) and list(hwf) == list(mine[2])
    assert all(np.array_equal(a, b) for a, b in zip(i_split, mine[3]))
    # PFM prior depths: our writer, the reference's reader
    d = np.random.RandomState(1).rand(5, 7).astype(np.float32)
    formats.write_pfm(str(tmp_path / "d.pfm"), d)
    got, scale = lb.read_pfm(str(tmp_path / "d.pfm"))
    assert np.array_equal(got, d) and scale == 1.0


def _dtu(tmp_path, views):
    from consistentnerf_b200 import formats
    rng = np.random.RandomState(0)
    w2c = np.tile(np.eye(4), (49, 1, 1))
    w2c[:, :3, 3] = rng.randn(49, 3) * 300
    w2c[:, :3, :3] = np.linalg.qr(rng.randn(49, 3, 3))[0]
    Kq = np.array([[180.0, 0, 80], [0, 180.0, 64], [0, 0, 1]])
    vd = {v: {"image": rng.rand(512, 640, 3), "depth": 2 + rng.rand(512, 640), "prior": (2 + rng.rand(512, 640)).astype(np.float32)} for v in views}
    formats.write_dtu_scan(str(tmp_path / "data" / "DTU"), "scan7", str(tmp_path), vd, w2c, Kq, 425.0, 2.5)
    return vd, w2c, Kq


def test_dtu_roundtrip_native(tmp_path):
    from consistentnerf_b200 import formats
    vd, w2c, Kq = _dtu(tmp_path, [3, 40])
    d = formats.load_dtu_scan(str(tmp_path / "data" / "DTU"), "scan7", str(tmp_path), [3, 40, 5])
    assert d["images"].shape == (3, 512, 640, 3) and d["K"][0, 0] == 720.0 and d["K"][0, 2] == 320.0
    assert np.abs(d["images"][0] - vd[3]["image"]).max() <= 0.5 / 255 + 1e-6 and d["images"][2].max() == 0.0
    np.testing.assert_allclose(d["depths"][1], vd[40]["depth"], rtol=1e-6)
    assert np.array_equal(d["priors"][0], vd[3]["prior"])
    np.testing.assert_allclose(d["bds"][0], [425 / 200, (425 + 2.5 * 192 * 1.06) / 200], rtol=1e-6)
    c2w = np.linalg.inv(w2c[3]); c2w[:3, 3] /= 200
    np.testing.assert_allclose(d["poses"][0][:, 0], c2w[:, 0], atol=1e-5)
    np.testing.assert_allclose(d["poses"][0][:, 1], -c2w[:, 1], atol=1e-5)      # OpenCV -> OpenGL axes (NP/load_dtu.py:203)
    intr, ext, nf = formats.read_dtu_cam(str(tmp_path / "data" / "DTU" / "Depths" / "Cameras" / "train" / "00000003_cam.txt"))
    np.testing.assert_allclose(ext, w2c[3], rtol=1e-6, atol=1e-4)


@needs_ref
def test_dtu_scan_is_read_by_the_reference_loader(tmp_path):
    from consistentnerf_b200 import formats
    ld = _import_ref("load_dtu")
    vd, w2c, Kq = _dtu(tmp_path, [25, 32])
    cwd = os.getcwd()
    os.chdir(tmp_path)                      # the loader reads the prior depths relative to the working directory
    try:
        imgs, poses, bds, render_poses, hwf, depths_cas, depths = ld.load_dtu_data("./data/DTU/scan7", train_view_num=3)
    finally:
        os.chdir(cwd)
    mine = formats.load_dtu_scan(str(tmp_path / "data" / "DTU"), "scan7", str(tmp_path), [25, 32, 0])
    assert imgs.shape == (49, 512, 640, 3) and hwf == [512, 640, 720.0] and render_poses.shape[1:] == (3, 4)
    for i, v in enumerate([25, 32, 0]):
        assert np.array_equal(imgs[v], mine["images"][i]) and np.array_equal(depths_cas[v], mine["priors"][i])
        assert np.array_equal(depths[v], mine["depths"][i])
        np.testing.assert_allclose(poses[v], mine["poses"][i], atol=1e-6)
    np.testing.assert_allclose([bds.min(), bds.max()], [mine["bds"].min(), mine["bds"].max()], rtol=1e-6)


def _llff(tmp_path, n=5, h=6, w=8, factor=8):
    from consistentnerf_b200 import formats
    rng = np.random.RandomState(2)
    imgs = (rng.rand(n, h, w, 3) * 255).astype(np.uint8)
    pb = np.zeros((n, 17))
    for i in range(n):
        p = np.concatenate([np.linalg.qr(rng.randn(3, 3))[0], rng.randn(3, 1), np.array([[h * factor], [w * factor], [100.0 * factor]])], 1)
        pb[i, :15], pb[i, 15:] = p.reshape(-1), [1.0 + rng.rand(), 5.0 + rng.rand()]
    formats.write_llff_scene(str(tmp_path / "fern"), imgs, pb, factor=factor)
    return imgs, pb


def test_llff_roundtrip_native(tmp_path):
    from consistentnerf_b200 import formats
    imgs, pb = _llff(tmp_path)
    poses, bds, loaded = formats.load_llff_scene(str(tmp_path / "fern"), factor=8)
    assert poses.shape == (3, 5, 5) and bds.shape == (2, 5) and loaded.shape == (6, 8, 3, 5)
    assert np.array_equal((loaded[..., 2] * 255 + 0.5).astype(np.uint8), imgs[2])
    assert poses[0, 4, 0] == 6 and poses[1, 4, 0] == 8 and poses[2, 4, 0] == 100.0      # hwf column: loaded shape, focal / factor
    np.testing.assert_allclose(bds[:, 1], pb[1, 15:])


@needs_ref
def test_llff_scene_is_read_by_the_reference_loader(tmp_path):
    from consistentnerf_b200 import formats
    ll = _import_ref("load_llff")
    imgs, pb = _llff(tmp_path)
    poses, bds, loaded, mono = ll._load_data(str(tmp_path / "fern"), factor=8)
    mine = formats.load_llff_scene(str(tmp_path / "fern"), factor=8)
    assert np.array_equal(poses, mine[0]) and np.array_equal(bds, mine[1]) and np.array_equal(loaded, mine[2])
    assert mono.shape[0] == 5


def test_pairs_file_loads_under_weights_only(tmp_path):
    from consistentnerf_b200 import formats
    formats.write_pairs(str(tmp_path / "configs" / "pairs.th"), {"dtu_train": np.array([25, 21, 33]), "dtu_val": [32, 24]})
    p = torch.load(str(tmp_path / "configs" / "pairs.th"))            # the scripts' own call (NP/run_nerf_view.py:942): default weights_only
    assert p["dtu_train"][:2] == [25, 21] and p["dtu_val"] == [32, 24]


def test_shims_cover_what_the_scripts_use(tmp_path):
    from consistentnerf_b200 import dropin
    dropin.install_io_stubs()
    import configargparse
    import imageio
    import lpips
    import pytorch_msssim
    import tensorboardX
    im = (np.random.RandomState(0).rand(5, 7, 4) * 255).astype(np.uint8)
    imageio.imwrite(str(tmp_path / "a.png"), im)
    assert np.array_equal(imageio.imread(str(tmp_path / "a.png")), im)
    assert np.array_equal(imageio.imread(str(tmp_path / "a.png"), ignoregamma=True)[..., :3], im[..., :3])
    p = configargparse.ArgumentParser()
    p.add_argument("--config", is_config_file=True, help="config file path")
    p.add_argument("--expname", type=str)
    p.add_argument("--N_rand", type=int, default=4096)
    p.add_argument("--white_bkgd", action="store_true")
    p.add_argument("--half_res", action="store_true")
    (tmp_path / "c.txt").write_text("expname = blender_paper_lego\nN_rand = 1024\nwhite_bkgd = True\nhalf_res = False\n# comment\n")
    a = p.parse_args(["--config", str(tmp_path / "c.txt"), "--N_rand", "77"])
    assert (a.expname, a.N_rand, a.white_bkgd, a.half_res) == ("blender_paper_lego", 77, True, False)
    x = torch.rand(1, 16, 16, 3, requires_grad=True)
    s = pytorch_msssim.ssim(x, x.detach(), data_range=1, size_average=False)
    assert s.shape == (1,) and abs(float(s) - 1.0) < 1e-5
    l = lpips.LPIPS(net="vgg").to(torch.device("cpu"))(x.permute(0, 3, 1, 2), x.permute(0, 3, 1, 2).detach())
    assert l.reshape(-1).shape == (1,) and float(l) == 0.0 and l.requires_grad
    w = tensorboardX.SummaryWriter(str(tmp_path / "runs"))
    w.add_scalar("loss", torch.tensor(0.5), 3)
    w.close()
    assert json.loads(open(tmp_path / "runs" / "scalars.jsonl").read())["value"] == 0.5


def test_metrics_file_with_cuda_tensor_repr(tmp_path):
    """the reference writes `LPIPS: tensor(0.1234, device='cuda:0', grad_fn=...)` (NP/run_nerf_view.py:2078-2087)."""
    from consistentnerf_b200 import formats
    (tmp_path / "metrics.txt").write_text("PSNR: 23.5\nSSIM: tensor(0.8125)\nLPIPS: tensor(0.1234, device='cuda:0', grad_fn=<MeanBackward0>)")
    m = formats.read_metrics(str(tmp_path / "metrics.txt"))
    assert m == {"PSNR": 23.5, "SSIM": 0.8125, "LPIPS": 0.1234}
    (tmp_path / "m2.txt").write_text("PSNR: 2.5e+01\nSSIM: -1.5E-3\nLPIPS: tensor(1.2500e-04, device='cuda:0')")
    assert formats.read_metrics(str(tmp_path / "m2.txt")) == {"PSNR": 25.0, "SSIM": -0.0015, "LPIPS": 0.000125}


@needs_ref
def test_llff_scene_through_the_reference_full_loader(tmp_path):
    """load_llff_data end to end (rescale by bd_factor, recenter, spiral render path, hold-out choice; NP/load_llff.py:288-356) on a
    forward-facing scene written by formats.write_llff_scene: 20 cameras on a small grid looking down -z, LLFF axis order."""
    from consistentnerf_b200 import formats
    ll = _import_ref("load_llff")
    rng = np.random.RandomState(5)
    n, h, w, factor, focal = 20, 6, 8, 8, 50.0
    imgs = (rng.rand(n, h, w, 3) * 255).astype(np.uint8)
    pb = np.zeros((n, 17))
    for i in range(n):
        pos = np.array([0.6 * (i % 5) / 4 - 0.3, 0.4 * (i // 5) / 3 - 0.2, 0.05 * rng.randn()])
        # camera axes in the world: right = +x, up = +y, back = +z (looking down -z); LLFF stores [down, right, back]
        R = np.stack([np.array([0, -1.0, 0]), np.array([1.0, 0, 0]), np.array([0, 0, 1.0])], 1)
        p = np.concatenate([R, pos[:, None], np.array([[h * factor], [w * factor], [focal * factor]])], 1)
        pb[i, :15], pb[i, 15:] = p.reshape(-1), [2.0 + 0.1 * rng.rand(), 8.0 + rng.rand()]
    formats.write_llff_scene(str(tmp_path / "fern"), imgs, pb, factor=factor)
    images, poses, bds, render_poses, i_test, mono = ll.load_llff_data(str(tmp_path / "fern"), factor, recenter=True, bd_factor=0.75, spherify=False)
    assert images.shape == (n, h, w, 3) and poses.shape == (n, 3, 5) and render_poses.shape == (60, 3, 4) and mono.shape[0] == n
    assert np.array_equal((images * 255 + 0.5).astype(np.uint8), imgs)
    sc = 1.0 / (pb[:, 15].min() * 0.75)
    np.testing.assert_allclose(bds, pb[:, 15:] * sc, rtol=1e-6)
    np.testing.assert_allclose(poses[:, :, 4], np.tile([h, w, focal], (n, 1)), rtol=1e-6)          # hwf: loaded shape, focal / factor
    # after the loader's axis swap the rotation is [right, up, back] = identity here, and recentring keeps it so
    np.testing.assert_allclose(poses[:, :3, :3], np.tile(np.eye(3), (n, 1, 1)), atol=1e-5)
    np.testing.assert_allclose(poses[:, :3, 3].mean(0), 0.0, atol=1e-5)                              # recentred around the mean camera
    assert 0 <= int(i_test) < n
    mine = formats.load_llff_scene(str(tmp_path / "fern"), factor=factor)
    assert np.array_equal(np.moveaxis(mine[2], -1, 0).astype(np.float32), images)


@needs_ref
def test_config1_reference_script_trains_on_the_cpu(tmp_path):
    """BASELINE config 1: the unmodified run_nerf.py -- Blender loader, 64 x 64, N_samples = 32, coarse only, no view directions -- runs its
    own train() on the CPU on a scene written by this package (formats + shims + twin harness; the product itself has no CPU path).
    10 iterations: plumbing, not quality."""
    sys.path.insert(0, ROOT)
    from oracle import twin
    if not os.path.isdir(twin.REF):
        pytest.skip("oracle/_ref not populated")
    root = str(tmp_path / "cfg1")
    twin.make_scene("blender", root, res=64)
    res = twin.run_arm_subprocess("ref", "blender", root, iters=10, eval_views=1, extra_args=twin.CONFIG1, device="cpu", timeout=900)
    assert "error" not in res, res
    assert res["device"] == "cpu" and res["patched"] == [] and res["checkpoints"] == []      # (run_nerf.py cannot checkpoint without a fine net)
    assert 5.0 < res["psnr"] < 60.0 and res["train_ms_per_iter"] > 0
    log = open(os.path.join(root, "log_blender_ref.txt")).read()
    assert "Loaded blender (5, 64, 64, 4)" in log and "TRAIN views are" in log            # 3 train + 1 val + 1 test views


def test_softmask_png_round_trip_through_the_imageio_shim(tmp_path):
    """--softmask masks (NP/run_nerf_view.py:1047-1054): written here, read the way the script does (imageio.imread / 255 > 0)."""
    from consistentnerf_b200 import formats, shims
    shims.install()
    import imageio
    rng = np.random.RandomState(2)
    mask = rng.rand(12, 16) > 0.7
    path = formats.write_softmask(str(tmp_path), "dtu", "scan21", 3, mask, top_k=30)
    assert path.endswith(os.path.join("Softmask", "dtu", "scan21", "iter_500", "softmask_0003_30per.png"))
    script_side = (imageio.imread(path).astype(np.float32) / 255.0).reshape(-1) > 0
    assert np.array_equal(script_side.reshape(12, 16), mask)
    assert np.array_equal(formats.read_softmask(str(tmp_path), "dtu", "scan21", 3), mask)


@needs_ref
def test_llff_reference_arm_trains_on_the_cpu(tmp_path):
    """BASELINE config 4's scene and harness without a GPU: twin.make_llff writes an LLFF-format fern stand-in (poses_bounds.npy,
    images_8/, prior depths in the loader's rescaled units, pairs); the UNMODIFIED run_nerf_view.py -- load_llff_data, hard-mask
    loop, NDC training (near 0 / far 1), checkpoint, held-out render -- runs on it on the CPU (reference arm only: `.cuda()` made a
    no-op by the harness; the product has no CPU path).  The hard masks cover most of each training view, i.e. the prior depths
    are consistent with the recentred poses."""
    import cv2
    sys.path.insert(0, ROOT)
    from oracle import twin
    if not os.path.isdir(twin.REF):
        pytest.skip("oracle/_ref not populated")
    root = str(tmp_path / "llff")
    twin.make_scene("llff", root)
    small = ["--use_viewdirs", "--N_samples", "32", "--N_importance", "16", "--N_rand", "512"]
    res = twin.run_arm_subprocess("ref", "llff", root, iters=4, eval_views=1, eval_res_div=3, extra_args=small, device="cpu", timeout=1500)
    assert "error" not in res, res
    assert res["device"] == "cpu" and res["patched"] == [] and res["checkpoints"] == ["000004.tar"]
    assert res["eval_hw"] == [126, 168] and 3.0 < res["psnr"] < 60.0
    log = open(os.path.join(root, "log_llff_ref.txt")).read()
    assert "Loaded llff (20, 378, 504, 3)" in log and "NEAR FAR 0.0 1.0" in log and "TRAIN views are [17, 2, 7, 6, 11, 1]" in log
    mdir = os.path.join(root, "logs", "twin_fern", "mask", "fern", "6view")
    cover = {v: float((cv2.imread(os.path.join(mdir, f"{v}_mask_6view.jpg"), 0) > 127).mean()) for v in range(20)}
    assert all(cover[v] > 0.6 for v in twin.FERN_TRAIN), cover
    assert all(cover[v] == 0.0 for v in range(20) if v not in twin.FERN_TRAIN)
