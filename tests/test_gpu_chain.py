"""SURVEY.md 8f-1 (first "next" row): the reference's whole per-frame chain — PreprocessProcessor (ROI crop +
INTER_AREA downscale) -> GrayscaleProcessor (BGR2GRAY) -> MagnificationProcessor — fused on the device behind
mc_chain_process, against the oracle's runChainOnce.  The "original" tap must be bit-exact (integer
arithmetic); the processed frame follows the magnification tolerances (<= 1 LSB for Laplace)."""
import numpy as np
import pytest

import lvm_b200 as L
from lvm_b200.synth import synth_frame
from oracle import livim_oracle as O
from common import make_cfgs, u8_diff

pytestmark = pytest.mark.gpu

CASES = [
    # (w, h, c, downscale, roi or None, grayscale)
    (640, 480, 3, 2, None, False),
    (640, 480, 3, 1, (0.25, 0.125, 0.5, 0.75), False),
    (641, 479, 3, 3, (0.1, 0.2, 0.77, 0.61), False),      # fractional INTER_AREA scale
    (640, 480, 3, 4, None, True),                          # downscale + gray
    (320, 240, 3, 1, None, True),                          # gray only
    (320, 240, 1, 2, (0.0, 0.0, 0.9, 0.9), True),          # gray input: GrayscaleProcessor is an identity
    (1920, 1080, 3, 8, None, False),
    (500, 300, 3, 7, (0.05, 0.05, 0.9, 0.9), True),
    (320, 240, 3, 1, None, False),                         # every front stage an identity
]


@pytest.mark.parametrize("w,h,c,down,roi,gray", CASES)
def test_chain_matches_oracle(w, h, c, down, roi, gray):
    cfg, ocfg = make_cfgs(O.MODE_LAPLACE, 20, 50.0, 0.4, 3.0, 20, 4)
    cfg.grayscale = ocfg.grayscale = gray
    for cc, PP in ((cfg, L.PreprocessParams), (ocfg, O.PreprocessParams)):
        cc.preprocess = PP(down, roi is not None, *(roi if roi else (0.0, 0.0, 1.0, 1.0)))
    chain, omag = L.ProcessingChainB200(0), O.MagnificationProcessor()
    for t in range(5):
        f = synth_frame(t, w, h, c)
        fr = L.Frame(image=f, seq=t)
        cur, orig = chain.run_chain_once(fr, cfg)
        ocur, oorig, cur_same, orig_same = O.run_chain_once(omag, f, ocfg)
        assert (orig is fr) == orig_same and (cur is fr) == cur_same, t
        if not orig_same:
            assert orig.image.shape == oorig.shape and np.array_equal(orig.image, oorig), t       # bit-exact tap
            assert (orig.width, orig.height) == (oorig.shape[1], oorig.shape[0])
        if not cur_same:
            assert cur.image.shape == ocur.shape, (cur.image.shape, ocur.shape)
            assert int(u8_diff(cur.image, ocur).max()) <= 1, t
            assert cur.seq == t


def test_chain_mode_none_returns_preprocessed_frame():
    """With magnification off the chain still crops / downsamples / greys (cur is the front stages' output)."""
    cfg = L.ProcessorConfig(grayscale=True, preprocess=L.PreprocessParams(2, False, 0, 0, 1, 1),
                            magnification=L.MagnificationParams(mode=L.MagnificationMode.NONE))
    ocfg = O.ProcessorConfig(grayscale=True, preprocess=O.PreprocessParams(2, False, 0, 0, 1, 1),
                             magnification=O.MagnificationParams(mode=O.MODE_NONE))
    chain, omag = L.ProcessingChainB200(0), O.MagnificationProcessor()
    f = synth_frame(0, 322, 200, 3)
    cur, orig = chain.run_chain_once(L.Frame(image=f), cfg)
    ocur, oorig, _, _ = O.run_chain_once(omag, f, ocfg)
    assert np.array_equal(cur.image, ocur) and np.array_equal(orig.image, oorig) and cur.format == "Gray8"


def test_roi_move_resets_temporal_state():
    """StructuralTracker compares PreprocessParams exactly (IProcessor.hpp:36-39): a moved ROI of equal size resets."""
    chain, omag = L.ProcessingChainB200(0), O.MagnificationProcessor()
    for t in range(8):
        roi = (0.1, 0.1, 0.5, 0.5) if t < 4 else (0.2, 0.1, 0.5, 0.5)
        cfg, ocfg = make_cfgs(O.MODE_LAPLACE, 20, 50.0, 0.4, 3.0, 0, 3)
        cfg.preprocess, ocfg.preprocess = L.PreprocessParams(1, True, *roi), O.PreprocessParams(1, True, *roi)
        f = synth_frame(t, 400, 300, 3)
        cur, _ = chain.run_chain_once(L.Frame(image=f), cfg)
        ocur, _, _, _ = O.run_chain_once(omag, f, ocfg)
        assert int(u8_diff(cur.image, ocur).max()) <= 1, t
