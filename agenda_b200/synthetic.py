"""Synthetic inputs for the post-process sweep (BASELINE.json configs[4], SURVEY.md §8d "Config 5")."""
from __future__ import annotations

import numpy as np


def synthetic_heatmaps(n: int, size: int = 512, seed: int = 0) -> np.ndarray:
    """fp32 [n,size,size]: sum of K in [0,40] Gaussian blobs (sigma in [4,16] px, amplitude U(0.2,1)) over a
    0.02*U(0,1) noise floor, rng = np.random.default_rng(seed) — vehicle-like components, not salt and pepper."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    out = np.empty((n, size, size), dtype=np.float32)
    for i in range(n):
        img = (0.02 * rng.random((size, size))).astype(np.float32)
        for _ in range(int(rng.integers(0, 41))):
            cx, cy = rng.uniform(0, size, 2)
            sig = rng.uniform(4, 16)
            amp = rng.uniform(0.2, 1.0)
            r = int(4 * sig) + 1
            x0, x1 = max(0, int(cx) - r), min(size, int(cx) + r + 1)
            y0, y1 = max(0, int(cy) - r), min(size, int(cy) + r + 1)
            img[y0:y1, x0:x1] += (amp * np.exp(-((xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2)
                                               / (2 * sig * sig))).astype(np.float32)
        out[i] = img
    return out
