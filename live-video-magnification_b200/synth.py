"""Deterministic synthetic clips: the same generator feeds the CUDA path, the oracle
and the CPU baseline.  numpy only — no oracle / cv2 dependency."""
from __future__ import annotations

import numpy as np


def synth_frame(t: int, w: int, h: int, channels: int = 3, fps: float = 30.0, seed: int = 0) -> np.ndarray:
    """Frame ``t`` of the synthetic clip: six oriented gratings (wavelengths 8..256 px) translating by
    0.3 px at 1.5 Hz (+0.1 px at 7 Hz), a 1.0 Hz grey-level pulse on a disc, and +-2 LSB uniform noise."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    shift = 0.3 * np.sin(2 * np.pi * 1.5 * t / fps) + 0.1 * np.sin(2 * np.pi * 7.0 * t / fps)
    img = np.zeros((h, w, 3), np.float64)
    for i, lam in enumerate((8, 16, 32, 64, 128, 256)):
        th = i * np.pi / 6.0
        u = (xx + shift) * np.cos(th) + (yy + 0.5 * shift) * np.sin(th)
        for c in range(3):
            img[:, :, c] += 14.0 * np.sin(2 * np.pi * u / lam + 0.7 * c + 0.3 * i)
    img += 128.0
    r = min(w, h) * 0.25
    disc = ((xx - w * 0.5) ** 2 + (yy - h * 0.5) ** 2) <= r * r
    img += (1.5 * np.sin(2 * np.pi * 1.0 * t / fps)) * disc[:, :, None]
    rng = np.random.default_rng(seed + t)
    img += rng.integers(-2, 3, size=(h, w, 3))
    out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if channels == 1:
        return np.ascontiguousarray(out[:, :, 0])
    return out


def synth_clip(n: int, w: int, h: int, channels: int = 3, fps: float = 30.0, seed: int = 0, start: int = 0):
    return [synth_frame(start + t, w, h, channels, fps, seed) for t in range(n)]
