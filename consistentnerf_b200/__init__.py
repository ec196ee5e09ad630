"""B200-native (sm_100a) per-ray volumetric rendering path of ConsistentNeRF.

Drop-in for the hot path of the reference's nerf-pytorch tree behind its own function surface:

    render / batchify_rays / render_rays / raw2outputs / run_network / batchify   (render.py)
    NeRF / Embedder / get_embedder / sample_pdf / get_rays / ndc_rays             (nerf.py)
    get_ref_rays / get_test_label / hard masks / masked and soft-weighted losses   (consistency.py)
    RayBank (device-resident batch sampler) / render_path (novel-view image loop) / StepLog (sync-free logging)   (pipeline.py)
    reference-format checkpoints / PFM depth maps / metrics.txt                   (formats.py)

All arithmetic runs in libcnerf.so (hand-written CUDA, C ABI in include/cnerf.h); there is no
PyTorch-eager or CPU fallback -- a missing library or a CPU tensor raises.
"""
from . import _lib  # noqa: F401
from .nerf import (NeRF, Embedder, get_embedder, sample_pdf, get_rays, get_rays_np, ndc_rays, img2mse, mse2psnr,
                   to8b)
from .render import batchify, run_network, batchify_rays, render, raw2outputs, render_rays, make_api
from .pipeline import RayBank, render_path, StepLog
from . import formats
from .consistency import (get_rays_ref, get_ref_rays, get_test_label, build_hard_masks, masked_img_loss,
                          masked_depth_loss, loss_scalars, img2mse_softmask, img2mse_depth_softmask, img2mse_softLpmask,
                          soft_depth_loss)

__version__ = "0.1.0"
