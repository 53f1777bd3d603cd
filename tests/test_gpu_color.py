"""Color mode (Gaussian pyramid + ideal FFT band-pass over a rolling window) parity vs the oracle.

Tolerance: pre-quantisation float output (input + colour image, [0,255] scale) within 1e-4 * 255 of the
oracle, u8 output <= 1 LSB (measured fp32 noise floor ~1e-6)."""
import numpy as np
import pytest

import lvm_b200 as L
from lvm_b200.synth import synth_frame
from oracle import livim_oracle as O
from common import make_cfgs, u8_diff

pytestmark = pytest.mark.gpu


def run(w, h, c, levels, n, fps, low=0.8, high=1.2, amp=100):
    cfg, ocfg = make_cfgs(O.MODE_COLOR, amp, 0.0, low, high, 0, levels, fps)
    proc, oproc = L.MagnificationProcessor(0), O.MagnificationProcessor()
    proc.set_option("keep_float_output", 1)
    worst_f, worst_u8, produced_any = 0.0, 0, 0
    for t in range(n):
        f = synth_frame(t, w, h, c, fps=fps)
        dbg = {}
        produced, out = proc.process_image(f, cfg)
        oprod, oout = oproc.process(f, ocfg, dbg)
        assert produced == oprod, t
        if not produced:
            assert out is f
            continue
        produced_any += 1
        got = proc.float_output(w, h, c)[0]
        ref = dbg["output_f32"] if c == 3 else dbg["output_f32"][..., None]
        worst_f = max(worst_f, float(np.abs(got - ref).max()) / 255.0)
        worst_u8 = max(worst_u8, int(u8_diff(out, oout).max()))
    assert produced_any == n - 1          # only the very first frame is a passthrough (window < 2 columns)
    return worst_f, worst_u8


@pytest.mark.parametrize("w,h,c,levels,n,fps", [
    (320, 240, 3, 3, 70, 30.0),     # window cap 64: warm-up lengths 2..64 (odd ones included) + 6 wrapped frames
    (320, 240, 1, 3, 24, 8.0),      # cap 16, wraps
    (250, 187, 3, 2, 20, 8.0),      # odd sizes -> bilinear resize after the pyrUp chain
    (130, 66, 3, 4, 20, 8.0),
])
def test_color_free_running(w, h, c, levels, n, fps):
    wf, wu = run(w, h, c, levels, n, fps)
    assert wf < 1e-4, wf
    assert wu <= 1, wu


def test_color_1080p_default_levels():
    wf, wu = run(1920, 1080, 3, 3, 5, 30.0)
    assert wf < 1e-4 and wu <= 1, (wf, wu)


def test_color_lanes_match_single():
    cfg, _ = make_cfgs(O.MODE_COLOR, 100, 0.0, 0.8, 1.2, 0, 3, 8.0)
    singles = [L.MagnificationProcessor(0) for _ in range(2)]
    multi = L.MagnificationProcessor(0, lanes=2)
    for t in range(20):
        frames = [synth_frame(t, 160, 120, 3, fps=8.0, seed=50 * k) for k in range(2)]
        outs = [p.process_image(f, cfg) for p, f in zip(singles, frames)]
        pm, mo = multi.process_image(np.stack(frames), cfg)
        assert pm == outs[0][0]
        if pm:
            for k in range(2):
                assert int(u8_diff(mo[k], outs[k][1]).max()) <= 1


def test_color_more_than_64_lanes():
    """Handles with more than 64 lanes: every lane's min/max slots must be initialised (round-1 advisor finding: k_mm_init
    covered only the first 256 slots, lane 69 of 70 came out 255 LSB off) — the last lanes equal a 1-lane handle."""
    w, h, lanes, n = 64, 48, 70, 5
    cfg, _ = make_cfgs(O.MODE_COLOR, 100, 0.0, 0.8, 1.2, 0, 2, 30.0)
    many = L.MagnificationProcessor(0, lanes=lanes)
    singles = {k: L.MagnificationProcessor(0) for k in (0, 64, 69)}
    for t in range(n):
        base = synth_frame(t, w, h, 3)
        frames = np.stack([np.roll(base, (k, 2 * k), axis=(0, 1)) for k in range(lanes)])
        produced, outs = many.process_image(frames, cfg)
        for k, p1 in singles.items():
            prod1, o1 = p1.process_image(frames[k], cfg)
            assert prod1 == produced
            if produced:
                assert np.array_equal(outs[k], o1), (t, k)


def test_color_framerate_changes_keep_window_semantics():
    """framerate is a live parameter: the window cap getOptimalBufferSize(int(fps)) grows (16 -> 64) and shrinks
    (64 -> 16, also while the ring is only partly filled) without a reset; the DFT length follows the window."""
    w, h, c, levels = 160, 120, 3, 2
    proc, oproc = L.MagnificationProcessor(0), O.MagnificationProcessor()
    fps_seq = [8.0] * 20 + [30.0] * 30 + [8.0] * 12 + [30.0] * 6 + [7.0] * 10
    for t, fps in enumerate(fps_seq):
        cfg, ocfg = make_cfgs(O.MODE_COLOR, 100, 0.0, 0.8, 1.2, 0, levels, fps)
        f = synth_frame(t, w, h, c, fps=8.0)
        produced, out = proc.process_image(f, cfg)
        oprod, oout = oproc.process(f, ocfg)
        assert produced == oprod, t
        if produced:
            assert int(u8_diff(out, oout).max()) <= 1, (t, fps, oproc.color.window.shape)
