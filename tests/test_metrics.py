"""metrics.py against independent direct-loop restatements of tf.image.psnr / tf.image.ssim (no TensorFlow here)."""
import math

import numpy as np
import torch

from video_prediction_b200 import metrics as M


def _ssim_loops(a, b, size=11, sigma=1.5, k1=0.01, k2=0.03):
    h, w, c = a.shape
    ax = np.arange(size) - (size - 1) / 2.0
    g = np.exp(-(ax[:, None] ** 2 + ax[None, :] ** 2) / (2 * sigma * sigma))
    g /= g.sum()
    c1, c2 = k1 ** 2, k2 ** 2
    vals = []
    for ch in range(c):
        acc = []
        for i in range(h - size + 1):
            for j in range(w - size + 1):
                x, y = a[i:i + size, j:j + size, ch], b[i:i + size, j:j + size, ch]
                mx, my = (g * x).sum(), (g * y).sum()
                sxx, syy, sxy = (g * x * x).sum() - mx * mx, (g * y * y).sum() - my * my, (g * x * y).sum() - mx * my
                acc.append(((2 * mx * my + c1) / (mx * mx + my * my + c1)) * ((2 * sxy + c2) / (sxx + syy + c2)))
        vals.append(np.mean(acc))
    return float(np.mean(vals))


def test_psnr_mse_ssim_match_direct_loops():
    rng = np.random.default_rng(0)
    a = rng.random((2, 3, 16, 20, 3))
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1)
    ta, tb = torch.tensor(a), torch.tensor(b)
    mse = M.mse(ta, tb).numpy()
    assert mse.shape == (2, 3) and np.allclose(mse, ((a - b) ** 2).mean(axis=(-3, -2, -1)))
    assert np.allclose(M.psnr(ta, tb).numpy(), -10 * np.log10(mse))
    s = M.ssim(ta, tb).numpy()
    for i in range(2):
        for j in range(3):
            assert abs(s[i, j] - _ssim_loops(a[i, j], b[i, j])) < 1e-9
    assert abs(float(M.ssim(ta, ta).min()) - 1.0) < 1e-12 and math.isinf(float(M.psnr(ta[0, 0], ta[0, 0])))
