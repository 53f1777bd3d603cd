#!/usr/bin/env python
"""Regenerates live-video-magnification_b200/csrc/lab_lut_s16.bin.

OpenCV's float cvtColor(BGR2Lab) does not evaluate the CIE formula: it interpolates a 33x33x33 int16
table (gamma + XYZ + cbrt baked in, values scaled by 2^14) with 4-bit fixed-point trilinear weights
(SURVEY.md §A.3).  The table is a constant of that algorithm, like a filter tap table; this script
recovers it exactly by converting the 33^3 lattice colours k/32 (where interpolation weights vanish)
and un-scaling: L*2^14/100, (a+128)*64, (b+128)*64 are exact integers (asserted).
Layout written: int16 little-endian [b][g][r][3] (b slowest), 33*33*33*3 values = 215622 bytes.
tests/test_oracle.py re-extracts the table at test time and checks the committed file against it.
"""
import os
import sys

import cv2
import numpy as np


def extract() -> np.ndarray:
    k = np.arange(33, dtype=np.float32) / np.float32(32)
    b, g, r = np.meshgrid(k, k, k, indexing="ij")
    img = np.stack([b, g, r], -1).reshape(33 * 33, 33, 3).astype(np.float32)
    lab = cv2.cvtColor(img, cv2.COLOR_BGR2Lab).reshape(33, 33, 33, 3).astype(np.float64)
    v = np.stack([lab[..., 0] * 16384 / 100, (lab[..., 1] + 128) * 64, (lab[..., 2] + 128) * 64], -1)
    assert np.abs(v - np.rint(v)).max() == 0.0, "lattice outputs are not exact table entries"
    return np.rint(v).astype("<i2")


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        os.path.dirname(os.path.abspath(__file__)), "..", "live-video-magnification_b200", "csrc",
        "lab_lut_s16.bin")
    lut = extract()
    lut.tofile(out)
    print("wrote", out, lut.shape, "cv2", cv2.__version__)
