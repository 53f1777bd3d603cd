#!/usr/bin/env python
"""Generates the committed golden vectors (tests/golden/*.npz) from the oracle.

The reference ships no tests or vectors for this path and cannot be built here, so
the vectors freeze what the cv2-based oracle (same OpenCV kernels, reference call order) produces, with
cv2's version recorded.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lvm_b200.synth import synth_frame  # noqa: E402
from oracle import livim_oracle as O  # noqa: E402

CASES = {
    # name: (mode, ui(amplification, wavelength, low, high, chroma, levels, fps), w, h, channels, frames)
    "laplace_color": (O.MODE_LAPLACE, (20, 50.0, 0.4, 3.0, 30, 4, 30.0), 96, 64, 3, 10),
    "laplace_gray": (O.MODE_LAPLACE, (20, 50.0, 0.4, 3.0, 0, 3, 30.0), 75, 50, 1, 10),
    "color_fft": (O.MODE_COLOR, (100, 0.0, 0.8, 1.2, 0, 2, 8.0), 96, 64, 3, 20),
    "riesz": (O.MODE_PHASE, (50, 50.0, 0.4, 3.0, 0, 3, 30.0), 96, 64, 3, 8),
}


def run_case(mode, ui, w, h, c, n):
    cfg = O.ProcessorConfig(magnification=O.to_params(mode, *ui))
    proc = O.MagnificationProcessor()
    frames, outs, produced = [], [], []
    for t in range(n):
        f = synth_frame(t, w, h, c, fps=ui[6])
        p, o = proc.process(f, cfg)
        frames.append(f)
        produced.append(bool(p))
        outs.append(o if p else np.zeros_like(f))
    extra = {}
    if mode == O.MODE_LAPLACE:
        extra["lowpassHi_1"] = proc.motion.lowpassHi[1]
        extra["lowpassLo_1"] = proc.motion.lowpassLo[1]
    return dict(frames=np.stack(frames), outputs=np.stack(outs), produced=np.array(produced), ui=np.array(ui, np.float64),
                mode=np.int32(mode), cv2_version=np.bytes_(cv2.__version__), **extra)


if __name__ == "__main__":
    cv2.setNumThreads(1)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, case in CASES.items():
        d = run_case(*case)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, d["frames"].shape, "produced", int(d["produced"].sum()), "cv2", cv2.__version__)
