"""Seeded synthetic rectified stereo pairs and binary masks (numpy only, deterministic).

Follows the recipe in SURVEY.md section 8(d): Gaussian-blurred uniform-noise texture, four
constant-disparity bands plus four random discs, left = right warped by the disparity plus
uniform noise in [-3, 3].  Frame i of a stream uses seed 1000 + i.
"""
from __future__ import annotations

import numpy as np


def _gauss_kernel(sigma: float) -> np.ndarray:
    r = max(1, int(np.ceil(3.0 * sigma)))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def gaussian_blur_u8(img: np.ndarray, sigma: float) -> np.ndarray:
    """Separable Gaussian blur with reflect-101 borders, float64 accumulate, round-half-even."""
    k = _gauss_kernel(sigma)
    r = len(k) // 2
    a = img.astype(np.float64)
    p = np.pad(a, ((0, 0), (r, r)), mode="reflect")
    a = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(len(k)))
    p = np.pad(a, ((r, r), (0, 0)), mode="reflect")
    a = sum(k[i] * p[i:i + img.shape[0], :] for i in range(len(k)))
    return np.clip(np.rint(a), 0, 255).astype(np.uint8)


def stereo_pair(width: int, height: int, ndisp: int, seed: int):
    """Returns (left, right, gt_disparity) as uint8/uint8/int32 arrays of shape (H, W)."""
    rng = np.random.default_rng(seed)
    W, H = width, height
    base = gaussian_blur_u8(rng.integers(0, 256, (H, W + ndisp + 8), dtype=np.uint8), 1.2)
    fresh = gaussian_blur_u8(rng.integers(0, 256, (H, W), dtype=np.uint8), 1.2)
    right = np.ascontiguousarray(base[:, :W])
    gt = np.zeros((H, W), np.int32)
    lo, hi = 3, max(4, ndisp - 8)
    band_d = rng.integers(lo, hi + 1, 4)
    for b in range(4):
        gt[b * H // 4:(b + 1) * H // 4, :] = band_d[b]
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(4):
        cx, cy = rng.integers(0, W), rng.integers(0, H)
        rad = rng.integers(8, max(9, H // 4))
        dd = rng.integers(1, max(2, ndisp - 2))
        gt[(xx - cx) ** 2 + (yy - cy) ** 2 <= rad * rad] = dd
    src = xx - gt
    ok = src >= 0
    left = np.where(ok, np.take_along_axis(right, np.clip(src, 0, W - 1), axis=1), fresh)
    noise = rng.integers(-3, 4, (H, W))
    left = np.clip(left.astype(np.int32) + noise, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left), right, gt


def binary_mask(width: int, height: int, seed: int) -> np.ndarray:
    """{0,255} mask like the inRange() output that feeds the filter plugin (estimator.cpp:43)."""
    rng = np.random.default_rng(seed)
    n = rng.integers(0, 256, (height, width), dtype=np.uint8)
    return np.where(gaussian_blur_u8(n, 4.0) > 128, 255, 0).astype(np.uint8)


def gray_image(width: int, height: int, seed: int) -> np.ndarray:
    """Generic gray-level image (exercises the filter on non-binary input)."""
    rng = np.random.default_rng(seed)
    return gaussian_blur_u8(rng.integers(0, 256, (height, width), dtype=np.uint8), 2.0)
