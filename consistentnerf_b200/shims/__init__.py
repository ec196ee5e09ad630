"""Stand-ins for the I/O-only third-party modules the reference scripts import.

``install()`` APPENDS this directory to ``sys.path``: a real installation of any of these modules, being earlier on the
path, always wins; the stand-in is only found when the module is absent (as in this image: no imageio, configargparse,
matplotlib, tensorboardX, ipdb, pytorch_msssim, lpips).

  imageio          imread / imwrite on cv2 (RGB(A) channel order, uint8 / uint16), mimwrite -> frames as PNGs
  configargparse   argparse + ``is_config_file`` arguments reading ``key = value`` files (the reference's configs/*.txt)
  tensorboardX     SummaryWriter that appends scalars to ``scalars.jsonl`` (one host sync per add_scalar, as the original)
  matplotlib       pyplot / cm: no-op figure calls, a grey->jet-like ``get_cmap``
  ipdb             set_trace() raises (a breakpoint in an unattended run is an error, not a hang)
  pytorch_msssim   ssim / ms_ssim: SSIM from global image statistics (no Gaussian window) -- a STAND-IN, not the metric
  lpips            LPIPS(...) whose call returns zeros -- a STAND-IN: the VGG weights are not available offline

The last two are not on the hot path ("out of hot-path scope"); train() of run_nerf_view.py calls them
unconditionally with small loss weights (NP/run_nerf_view.py:1701-1728), so they must exist for the script to run.  Both twin
arms (reference eager and this package) see the same stand-ins.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def install() -> str:
    if HERE not in sys.path:
        sys.path.append(HERE)
    return HERE
