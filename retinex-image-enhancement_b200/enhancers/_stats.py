"""Host-side arithmetic on the 256-bin gray histogram (a3).

The reference (enhancers/adaptive_params.py:52-66) runs five NumPy passes over the u8 gray image;
every one of them is an exact function of the histogram the GPU kernel returns:
mean = sum(k*h)/N, population std = sqrt(sum(h*(k-mean)^2)/N), ratios = partial sums / N.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import native


def brightness_histogram(x: torch.Tensor) -> np.ndarray:
    """[N,3,H,W] f32 CUDA -> [N,256] int64 host histogram (one 1 KB D2H per image)."""
    return native.brightness_hist(x).cpu().numpy().astype(np.int64)


def features_from_histogram(hist: np.ndarray):
    out = []
    k = np.arange(256, dtype=np.float64)
    for h in np.atleast_2d(hist):
        n = int(h.sum())
        mean = float((h * k).sum()) / n
        var = float((h * (k - mean) ** 2).sum()) / n
        out.append({
            "mean_brightness": mean / 255.0,
            "brightness_std": float(np.sqrt(var)) / 255.0,
            "dark_pixel_ratio": float(h[:50].sum()) / n,       # gray < 50
            "mid_pixel_ratio": float(h[50:201].sum()) / n,     # 50 <= gray <= 200
            "bright_pixel_ratio": float(h[201:].sum()) / n,    # gray > 200
        })
    return out
