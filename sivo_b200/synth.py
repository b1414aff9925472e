"""Seeded synthetic 1242x375 stereo frames: band-limited noise plus high-contrast
rectangles and lines so FAST fires in most cells; the right image is the left one shifted by a
per-row-band disparity.  There is no KITTI data in the build or bench containers."""
from __future__ import annotations

import numpy as np

RAW_W, RAW_H = 1242, 375


def _blur(a: np.ndarray, sigma: float) -> np.ndarray:
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(a, ((0, 0), (r, r)), mode="reflect")
    a = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(2 * r + 1))
    p = np.pad(a, ((r, r), (0, 0)), mode="reflect")
    return sum(k[i] * p[i:i + a.shape[0]] for i in range(2 * r + 1))


def stereo_frame(frame_idx: int = 0, w: int = RAW_W, h: int = RAW_H):
    """Returns (left_bgr u8 [h,w,3], right_bgr u8 [h,w,3])."""
    rng = np.random.default_rng(1000 + frame_idx)
    base = _blur(rng.uniform(0, 255, size=(h, w)), 3.0)
    base = (base - base.min()) / (base.max() - base.min()) * 255.0
    img = np.repeat(base[:, :, None], 3, axis=2)
    img += rng.normal(0, 6.0, size=(1, 1, 3))  # slight colour cast per channel
    for _ in range(200):
        x0 = int(rng.integers(0, w - 8))
        y0 = int(rng.integers(0, h - 8))
        col = rng.uniform(0, 255, size=3)
        if rng.random() < 0.6:
            rw, rh = int(rng.integers(6, 60)), int(rng.integers(6, 40))
            img[y0:y0 + rh, x0:x0 + rw] = col
        else:
            ln = int(rng.integers(20, 120))
            t = int(rng.integers(1, 4))
            if rng.random() < 0.5:
                img[y0:y0 + t, x0:x0 + ln] = col
            else:
                img[y0:y0 + ln, x0:x0 + t] = col
    left = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    right = np.empty_like(left)
    band = 25
    for y0 in range(0, h, band):
        d = int(rng.integers(5, 61))
        rows = left[y0:y0 + band]
        right[y0:y0 + band, :w - d] = rows[:, d:]
        right[y0:y0 + band, w - d:] = rows[:, -1:, :]
    return left, right


def bgr_to_gray(bgr: np.ndarray) -> np.ndarray:
    """cv::cvtColor(BGR2GRAY) for 8-bit as cv2 4.13 computes it: (B*3735 + G*19235 + R*9798 + 16384) >> 15
    (the reference converts in Tracking::GrabImageStereo, src/orbslam/Tracking.cc:187-194)."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)
