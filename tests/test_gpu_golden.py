"""CUDA path vs the committed golden vectors (tests/golden/*.npz) — no oracle import needed on the box."""
import os

import numpy as np
import pytest

import lvm_b200 as L
from common import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["laplace_color", "laplace_gray", "color_fft", "riesz"])
def test_cuda_matches_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    ui = g["ui"]
    vals = L.MagUiValues(L.MagnificationMode(int(g["mode"])), int(ui[0]), float(ui[1]), float(ui[2]), float(ui[3]),
                         int(ui[4]), int(ui[5]), float(ui[6]))
    cfg = L.ProcessorConfig(magnification=L.toParams(vals))
    proc = L.MagnificationProcessor(0)
    for t, f in enumerate(g["frames"]):
        produced, out = proc.process_image(np.ascontiguousarray(f), cfg)
        assert produced == bool(g["produced"][t]), t
        if not produced:
            continue
        d = np.abs(out.astype(np.int32) - g["outputs"][t].astype(np.int32))
        if name == "riesz":
            assert d.max() <= 3 and (d == 0).mean() >= 0.995, (t, d.max())
        else:
            assert d.max() <= 1, (t, d.max())
    if name == "laplace_color":
        hi = proc.get_state("lowpassHi", 1)[0]
        assert np.abs(np.moveaxis(hi, 0, 2) - g["lowpassHi_1"]).max() < 1e-3
