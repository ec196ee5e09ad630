"""imageio stand-in on OpenCV: the handful of calls the reference's loaders and scripts make."""
import os

import cv2
import numpy as np

__version__ = "0-cnerf-shim"


def imread(uri, ignoregamma=None, **_kw):
    """-> ndarray [H,W], [H,W,3] (RGB) or [H,W,4] (RGBA), dtype as stored (uint8 / uint16)."""
    img = cv2.imread(str(uri), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(f"imageio(shim).imread: cannot read {uri}")
    if img.ndim == 3 and img.shape[2] == 3:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    elif img.ndim == 3 and img.shape[2] == 4:
        img = cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA)
    return img


def imwrite(uri, im, **_kw):
    im = np.asarray(im)
    if im.dtype == bool:
        im = im.astype(np.uint8) * 255
    if im.dtype not in (np.uint8, np.uint16):
        im = np.clip(im, 0, 255).astype(np.uint8)
    if im.ndim == 3 and im.shape[2] == 3:
        im = cv2.cvtColor(im, cv2.COLOR_RGB2BGR)
    elif im.ndim == 3 and im.shape[2] == 4:
        im = cv2.cvtColor(im, cv2.COLOR_RGBA2BGRA)
    if not cv2.imwrite(str(uri), im):
        raise IOError(f"imageio(shim).imwrite: cannot write {uri}")


imsave = imwrite


def mimwrite(uri, ims, fps=30, quality=8, **_kw):
    """No video encoder here: the frames are written as <uri>.frames/%04d.png."""
    d = str(uri) + ".frames"
    os.makedirs(d, exist_ok=True)
    for i, im in enumerate(ims):
        imwrite(os.path.join(d, f"{i:04d}.png"), im)


mimsave = mimwrite
