"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

import lvm_b200 as L
from oracle import livim_oracle as O

GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def make_cfgs(mode, amplification, wavelength, low, high, chroma, levels, fps=30.0):
    """-> (product ProcessorConfig, oracle ProcessorConfig) from the same UI values."""
    ui = L.MagUiValues(mode=L.MagnificationMode(mode), amplification=amplification, wavelength=wavelength, low=low,
                       high=high, chroma=chroma, levels=levels, captureFps=fps)
    cfg = L.ProcessorConfig(magnification=L.toParams(ui))
    ocfg = O.ProcessorConfig(magnification=O.to_params(mode, amplification, wavelength, low, high, chroma, levels, fps))
    return cfg, ocfg


def planar(m):
    """oracle HxWxC (or HxW) f32 -> [C][H][W]"""
    m = np.asarray(m)
    return m[None] if m.ndim == 2 else np.ascontiguousarray(np.moveaxis(m, 2, 0))


def u8_diff(a, b):
    return np.abs(a.astype(np.int32) - b.astype(np.int32))
