#!/usr/bin/env python
"""Which summation does cv2.filter2D use for the 9x9 f32 Riesz kernels on THIS host?  Raster-order FMA in the vectorised
columns x < floor(w / V) * V and multiply-then-add in the scalar tail, with V the vector width of the dispatched ISA
(8 lanes with AVX2, 16 with AVX-512).  Prints V — the reference's band planes, and through acos + amplification its Phase
output, depend on it in the tail columns."""
import json
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import livim_oracle as O   # noqa: E402  (the tap table only)


def direct(img, K, fma):
    pad = cv2.copyMakeBorder(img, 4, 4, 4, 4, cv2.BORDER_REFLECT_101)
    h, w = img.shape
    acc = np.zeros((h, w), np.float32)
    for ky in range(9):
        for kx in range(9):
            c = K[ky, kx]
            if c == 0:
                continue
            v = pad[ky:ky + h, kx:kx + w]
            acc = (np.float64(c) * v.astype(np.float64) + acc.astype(np.float64)).astype(np.float32) if fma else (acc + (np.float32(c) * v)).astype(np.float32)
    return acc


def main():
    K = np.asarray(O.HIGHPASS_9x9, np.float32)
    rng = np.random.default_rng(0)
    out = {}
    for w in (60, 71, 120, 135):
        img = (rng.random((12, w)) * 100).astype(np.float32)
        ref = cv2.filter2D(img, -1, K, borderType=cv2.BORDER_REFLECT_101)
        f, m = direct(img, K, True), direct(img, K, False)
        cols = np.where((ref != f).any(axis=0))[0]
        first = int(cols[0]) if len(cols) else w
        out[str(w)] = {"first_non_fma_column": first, "tail_is_mul_add": bool((ref[:, first:] == m[:, first:]).all()),
                       "head_is_fma": bool((ref[:, :first] == f[:, :first]).all())}
    model = "?"
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    print(json.dumps({"cpu": model, "cv2": cv2.__version__, "filter2D_9x9_f32": out}))


if __name__ == "__main__":
    main()
