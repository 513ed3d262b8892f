"""Evaluation metrics of the reference (video_prediction/metrics.py:5-14): mse, psnr (tf.image.psnr, max_val 1) and ssim
(tf.image.ssim, max_val 1: 11x11 Gaussian window sigma 1.5, k1 0.01, k2 0.03, VALID filtering, mean over the filtered
positions and the channels).  lpips / vgg distances need network weights that cannot be downloaded here and are not
provided.  Evaluation is outside the training hot path: these run as a handful of torch ops on whatever
device the frames live on; inputs [..., H, W, C] in [0, 1], outputs [...]."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def mse(a, b):
    return ((a - b) ** 2).mean(dim=(-3, -2, -1))


def psnr(a, b, max_val=1.0):
    return 20.0 * math.log10(max_val) - 10.0 * torch.log10(mse(a, b))


def _gaussian_window(size=11, sigma=1.5, dtype=torch.float32, device=None):
    # tf.image.ssim's _fspecial_gauss: softmax of -(x^2 + y^2) / (2 sigma^2) over the window
    c = torch.arange(size, dtype=dtype, device=device) - (size - 1) / 2.0
    g = -(c ** 2) / (2.0 * sigma * sigma)
    g2 = g[:, None] + g[None, :]
    return torch.softmax(g2.reshape(-1), dim=0).reshape(size, size)


def ssim(a, b, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    lead = a.shape[:-3]
    h, w, c = a.shape[-3:]
    x = a.reshape(-1, h, w, c).permute(0, 3, 1, 2).reshape(-1, 1, h, w)
    y = b.reshape(-1, h, w, c).permute(0, 3, 1, 2).reshape(-1, 1, h, w)
    win = _gaussian_window(filter_size, filter_sigma, a.dtype, a.device)[None, None]
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    mx, my = F.conv2d(x, win), F.conv2d(y, win)
    # tf.image.ssim computes the luminance term from the means and the contrast-structure term from the filtered moments
    sxx = F.conv2d(x * x, win) - mx * mx
    syy = F.conv2d(y * y, win) - my * my
    sxy = F.conv2d(x * y, win) - mx * my
    lum = (2 * mx * my + c1) / (mx * mx + my * my + c1)
    cs = (2 * sxy + c2) / (sxx + syy + c2)
    per_channel = (lum * cs).mean(dim=(-2, -1)).reshape(-1, c)
    return per_channel.mean(dim=-1).reshape(lead)


METRIC_FNS = (('psnr', psnr), ('mse', mse), ('ssim', ssim))
