#!/usr/bin/env python
"""Generates the committed golden vectors (tests/golden/*.npz) by running the REFERENCE ITSELF.

The reference ships no tests or vectors for this path.  The vectors are the outputs of the
reference's own sources compiled in place (oracle/_ref/_livim_ref: /root/reference/src/processing/** against the
cvshim facade, pixel operations executed by the real OpenCV kernels in cv2 — oracle/build_ref.py), with cv2's
version recorded; the script asserts that the oracle restatement reproduces every one of them bit for bit
before writing.  Needs /root/reference (this container).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lvm_b200.synth import synth_frame  # noqa: E402
from oracle import livim_oracle as O  # noqa: E402
from oracle import livim_ref  # noqa: E402

CASES = {
    # name: (mode, ui(amplification, wavelength, low, high, chroma, levels, fps), w, h, channels, frames)
    "laplace_color": (O.MODE_LAPLACE, (20, 50.0, 0.4, 3.0, 30, 4, 30.0), 96, 64, 3, 10),
    "laplace_gray": (O.MODE_LAPLACE, (20, 50.0, 0.4, 3.0, 0, 3, 30.0), 75, 50, 1, 10),
    "color_fft": (O.MODE_COLOR, (100, 0.0, 0.8, 1.2, 0, 2, 8.0), 96, 64, 3, 20),
    "riesz": (O.MODE_PHASE, (50, 50.0, 0.4, 3.0, 0, 3, 30.0), 96, 64, 3, 8),
}


def run_case(R, mode, ui, w, h, c, n):
    ocfg = O.ProcessorConfig(magnification=O.to_params(mode, *ui))
    rparams = livim_ref.to_ref_params(R, ocfg.magnification)
    levels = min(max(ocfg.magnification.levels, 1), R.calculateMaxLevels(w, h))
    core, oproc = R.Core(), O.MagnificationProcessor()     # Core = magcore::magnify* with readable state
    frames, outs, produced = [], [], []
    for t in range(n):
        f = synth_frame(t, w, h, c, fps=ui[6])
        p, o = core.run(mode, f, rparams, levels)
        po, oo = oproc.process(f, ocfg)
        assert p == po and (not p or np.array_equal(o, oo)), f"oracle != reference at frame {t}"
        frames.append(f)
        produced.append(bool(p))
        outs.append(o if p else np.zeros_like(f))
    extra = {}
    if mode == O.MODE_LAPLACE:
        hi, lo = core.motion_state()
        assert np.array_equal(hi[1], oproc.motion.lowpassHi[1]) and np.array_equal(lo[1], oproc.motion.lowpassLo[1])
        extra["lowpassHi_1"], extra["lowpassLo_1"] = hi[1], lo[1]
    return dict(frames=np.stack(frames), outputs=np.stack(outs), produced=np.array(produced), ui=np.array(ui, np.float64),
                mode=np.int32(mode), cv2_version=np.bytes_(cv2.__version__),
                source=np.bytes_("reference sources compiled in place (oracle/_ref/_livim_ref)"), **extra)


if __name__ == "__main__":
    cv2.setNumThreads(1)
    R = livim_ref.load()
    assert R is not None, "needs /root/reference to build oracle/_ref/_livim_ref"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, case in CASES.items():
        d = run_case(R, *case)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, d["frames"].shape, "produced", int(d["produced"].sum()), "cv2", cv2.__version__)
