"""pytorch_msssim STAND-IN (not the metric): SSIM from per-image global statistics, no Gaussian window, no multi-scale pyramid.

Differentiable, in [-1, 1], 1 for identical inputs -- enough for train() of NP/run_nerf_view.py (:1701, weight 0.005 on 16x16
patches) and its test-set report to run; numbers written to metrics.txt under this stand-in are NOT comparable with SSIM."""
import torch


def _global_ssim(X, Y, data_range):
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    X, Y = X.flatten(1).to(torch.float32), Y.flatten(1).to(torch.float32)
    mx, my = X.mean(1), Y.mean(1)
    vx, vy = X.var(1, unbiased=False), Y.var(1, unbiased=False)
    cov = ((X - mx[:, None]) * (Y - my[:, None])).mean(1)
    return ((2 * mx * my + C1) * (2 * cov + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2))


def ssim(X, Y, data_range=255, size_average=True, **_kw):
    s = _global_ssim(X, Y, float(data_range))
    return s.mean() if size_average else s


def ms_ssim(X, Y, data_range=255, size_average=True, **_kw):
    return ssim(X, Y, data_range=data_range, size_average=size_average)
