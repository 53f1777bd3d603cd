"""Motion (Phase / Riesz) parity vs the oracle.

Two tiers (SURVEY.md §A.7 — acos near 1 and the unbounded phase integrator make free-running parity
chaotic for any implementation that is not bit-identical to OpenCV's 9x9 filter summation order):
  tier 1  teacher-forced single step: oracle state of frame t-1 injected, frame t compared:
          float output (BGR in [0,1], before quantisation) max-abs < 1e-4, u8 <= 1 LSB
  tier 2  free-running over >= 32 frames: u8 <= 3 LSB, >= 99.5 % of samples identical,
          p99.9 of the float error < 3e-3."""
import numpy as np
import pytest

import lvm_b200 as L
from lvm_b200.synth import synth_frame
from oracle import livim_oracle as O
from common import make_cfgs, u8_diff

pytestmark = pytest.mark.gpu

STATE = [("old.lowpass", lambda st, i: st.old.levels[i].lowpass), ("old.rx", lambda st, i: st.old.levels[i].rx),
         ("old.ry", lambda st, i: st.old.levels[i].ry),
         ("phase.c", lambda st, i: st.lo.phase[i][0]), ("phase.s", lambda st, i: st.lo.phase[i][1]),
         ("lo.r0.c", lambda st, i: st.lo.reg0[i][0]), ("lo.r0.s", lambda st, i: st.lo.reg0[i][1]),
         ("lo.r1.c", lambda st, i: st.lo.reg1[i][0]), ("lo.r1.s", lambda st, i: st.lo.reg1[i][1]),
         ("hi.r0.c", lambda st, i: st.hi.reg0[i][0]), ("hi.r0.s", lambda st, i: st.hi.reg0[i][1]),
         ("hi.r1.c", lambda st, i: st.hi.reg1[i][0]), ("hi.r1.s", lambda st, i: st.hi.reg1[i][1])]


def inject(proc, ost, levels):
    for i in range(levels - 1):
        for name, get in STATE:
            proc.set_state(name, i, np.ascontiguousarray(get(ost, i))[None, None])


@pytest.mark.parametrize("w,h,levels", [(320, 240, 4), (250, 187, 3), (96, 64, 2)])
def test_teacher_forced_single_step(w, h, levels):
    cfg, ocfg = make_cfgs(O.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, levels)
    proc, oproc = L.MagnificationProcessor(0), O.MagnificationProcessor()
    proc.set_option("keep_float_output", 1)
    worst = 0.0
    for t in range(8):
        f = synth_frame(t, w, h, 3)
        if t >= 2:
            inject(proc, oproc.riesz, levels)
        dbg = {}
        produced, out = proc.process_image(f, cfg)
        oprod, oout = oproc.process(f, ocfg, dbg)
        assert produced == oprod == (t >= 1)
        if not produced:
            continue
        err = float(np.abs(proc.float_output(w, h, 3)[0] - dbg["output_bgr_f32"]).max())
        worst = max(worst, err)
        assert err < 1e-4, (t, err)
        assert int(u8_diff(out, oout).max()) <= 1, t
        # the updated state must track the oracle too.  Pyramid planes (Lab scale): 1e-4 absolute.  Phase
        # accumulators / IIR registers: p99.9 < 1e-4 rad.  Isolated pixels may jump: the reference's arcCos
        # returns +-1.0 *radians* when q_r/|q| rounds 1 ulp outside [-1, 1] (RieszPyramid.cpp:15-18),
        # and acos(1 - eps) ~ sqrt(2 eps), so a last-bit difference in the 9x9 sums can move one pixel's
        # phase by up to 1 rad; such pixels are bounded in number (<= 1e-4 of the plane), not in size.
        for i in range(levels - 1):
            for name, get in STATE:
                e = np.abs(proc.get_state(name, i)[0, 0] - get(oproc.riesz, i))
                if name.startswith("old."):
                    assert float(e.max()) < 1e-4, (t, i, name, float(e.max()))
                else:
                    assert float(np.quantile(e, 0.999)) < 1e-4 and float((e > 1e-3).mean()) <= 1e-4, \
                        (t, i, name, float(e.max()), float(np.quantile(e, 0.999)), float((e > 1e-3).mean()))


def test_free_running_32_frames():
    w, h, levels = 320, 240, 4
    cfg, ocfg = make_cfgs(O.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, levels)
    proc, oproc = L.MagnificationProcessor(0), O.MagnificationProcessor()
    proc.set_option("keep_float_output", 1)
    for t in range(34):
        f = synth_frame(t, w, h, 3)
        dbg = {}
        produced, out = proc.process_image(f, cfg)
        oprod, oout = oproc.process(f, ocfg, dbg)
        assert produced == oprod
        if not produced:
            continue
        d8 = u8_diff(out, oout)
        ferr = np.abs(proc.float_output(w, h, 3)[0] - dbg["output_bgr_f32"])
        assert int(d8.max()) <= 3, (t, int(d8.max()))
        assert float((d8 == 0).mean()) >= 0.995, (t, float((d8 == 0).mean()))
        assert float(np.quantile(ferr, 0.999)) < 3e-3, t


def test_gray_and_first_frame_passthrough():
    proc = L.MagnificationProcessor(0)
    cfg, _ = make_cfgs(O.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, 4)
    g = synth_frame(0, 160, 120, 1)
    for t in range(3):
        produced, out = proc.process_image(g, cfg)       # gray input: silent passthrough (MagnifyCore.hpp:212)
        assert not produced and out is g
    f = synth_frame(0, 160, 120, 3)
    produced, out = proc.process_image(f, cfg)           # channel change -> reset -> first frame passthrough
    assert not produced and out is f
    produced, out = proc.process_image(synth_frame(1, 160, 120, 3), cfg)
    assert produced


def test_cutoff_change_rebuilds_old_and_zeroes_filters():
    w, h, levels = 200, 150, 3
    proc, oproc = L.MagnificationProcessor(0), O.MagnificationProcessor()
    seq = [(0.4, 3.0)] * 5 + [(0.8, 3.0)] * 4 + [(0.8, 2.0)] * 4
    for t, (lo, hi) in enumerate(seq):
        cfg, ocfg = make_cfgs(O.MODE_PHASE, 30, 50.0, lo, hi, 0, levels)
        f = synth_frame(t, w, h, 3)
        produced, out = proc.process_image(f, cfg)
        oprod, oout = oproc.process(f, ocfg)
        assert produced == oprod
        if produced:
            d8 = u8_diff(out, oout)
            assert int(d8.max()) <= 3 and float((d8 == 0).mean()) >= 0.995, (t, int(d8.max()))


def test_config3_1080p():
    w, h, levels = 1920, 1080, 6
    cfg, ocfg = make_cfgs(O.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, levels)
    proc, oproc = L.MagnificationProcessor(0), O.MagnificationProcessor()
    for t in range(4):
        f = synth_frame(t, w, h, 3)
        produced, out = proc.process_image(f, cfg)
        oprod, oout = oproc.process(f, ocfg)
        assert produced == oprod
        if produced:
            d8 = u8_diff(out, oout)
            assert int(d8.max()) <= 3 and float((d8 == 0).mean()) >= 0.995, (t, int(d8.max()), float((d8 == 0).mean()))


def test_riesz_lanes_match_single():
    """A 2-lane handle (two independent streams stepped in lock-step) equals two 1-lane handles bit for bit."""
    cfg, _ = make_cfgs(O.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, 3)
    singles = [L.MagnificationProcessor(0) for _ in range(2)]
    multi = L.MagnificationProcessor(0, lanes=2)
    for t in range(6):
        frames = [synth_frame(t, 160, 120, 3, seed=70 * k) for k in range(2)]
        outs = [p.process_image(f, cfg) for p, f in zip(singles, frames)]
        pm, mo = multi.process_image(np.stack(frames), cfg)
        assert pm == outs[0][0] == outs[1][0]
        if pm:
            for k in range(2):
                assert np.array_equal(mo[k], outs[k][1]), (t, k)


def test_flat_regions_nan_semantics_match_the_reference():
    """Letterbox bars / black or flat frames: the blurred amplitude is exactly 0 there, the reference divides 0/0
    (RieszPyramid.cpp:125-126), the NaN survives cosSin and the collapse, and cv::cvtColor(Lab2BGR) clips it to
    1.0 — the bars come out WHITE.  The device path must reproduce exactly that (same NaN region, same mapping)."""
    w, h, levels = 192, 128, 4
    cfg, ocfg = make_cfgs(O.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, levels)
    rng = np.random.default_rng(3)

    def letterbox(t):
        f = np.zeros((h, w, 3), np.uint8)
        f[32:96] = synth_frame(t, w, 64, 3)
        return f

    clips = {"letterbox": [letterbox(t) for t in range(5)],
             "black": [np.zeros((h, w, 3), np.uint8) for _ in range(3)],
             "flat+noise block": [np.full((h, w, 3), 128, np.uint8) for _ in range(3)]}
    for f in clips["flat+noise block"]:
        f[40:80, 50:120] = rng.integers(90, 170, size=(40, 70, 3))
    for name, frames in clips.items():
        proc, oproc = L.MagnificationProcessor(0), O.MagnificationProcessor()
        for t, f in enumerate(frames):
            produced, out = proc.process_image(f, cfg)
            oprod, oout = oproc.process(f, ocfg)
            assert produced == oprod, (name, t)
            if produced:
                d = u8_diff(out, oout)
                assert int(d.max()) <= 3 and float((d == 0).mean()) >= 0.995, (name, t, int(d.max()), float((d == 0).mean()))
                if name == "letterbox":
                    assert (oout[:8] == 255).all() and (out[:8] == 255).all()      # the bars are white in both


def test_tma_staged_riesz_tiles_equal_ldg_staging():
    """The 9x9 analysis and collapse kernels stage their interior tiles with one cp.async.bulk.tensor copy each (option
    use_tma, default on); use_tma = 0 selects 128-bit loads.  Bit-identical u8 frames and state planes, on a size with
    interior and border tiles at several levels."""
    w, h, levels = 400, 300, 4
    cfg, _ = make_cfgs(O.MODE_PHASE, 40, 50.0, 0.4, 3.0, 0, levels, 30.0)
    a, b = L.MagnificationProcessor(0), L.MagnificationProcessor(0)
    b.set_option("use_tma", 0)
    for t in range(5):
        f = synth_frame(t, w, h, 3)
        pa, oa = a.process_image(f, cfg)
        pb, ob = b.process_image(f, cfg)
        assert pa == pb
        if pa:
            assert np.array_equal(oa, ob), t
    for lvl in range(levels - 1):
        for name in ("old.lowpass", "old.rx", "phase.c", "lo.r0.c"):
            assert np.array_equal(a.get_state(name, lvl), b.get_state(name, lvl)), (name, lvl)


@pytest.mark.parametrize("w,h,levels", [(160, 120, 4), (71, 76, 4), (135, 90, 4), (322, 241, 5)])
def test_band_planes_and_riesz_pair_are_bit_identical_to_the_reference(w, h, levels):
    """cv::filter2D accumulates the 9x9 / 1x5 / 5x1 taps in raster order with one FMA per tap in its vectorised columns
    (x < (w & ~7)) and with multiply-then-add in the scalar tail columns; the device kernels follow that rule column by
    column (mc_riesz.cu::f2d).  The band planes of every level and the Riesz pair therefore equal the reference's BIT FOR
    BIT, for widths that are not multiples of 8 too — which is what keeps the ill-conditioned acos(q_real / |q|) from
    amplifying last-ulp differences (before this rule a 71-wide frame differed in 3.3 % of its u8 samples, now in none
    on the CPU emulation)."""
    cfg, ocfg = make_cfgs(O.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, levels, 30.0)
    proc, op = L.MagnificationProcessor(0), O.MagnificationProcessor()
    for t in range(3):
        f = synth_frame(t, w, h, 3)
        proc.process_image(f, cfg)
        op.process(f, ocfg)
    for lvl in range(levels - 1):
        ref = op.riesz.old.levels[lvl]
        for name, plane in (("old.lowpass", ref.lowpass), ("old.rx", ref.rx), ("old.ry", ref.ry)):
            got = proc.get_state(name, lvl)[0, 0]
            assert np.array_equal(got, np.asarray(plane)), (name, lvl, float(np.abs(got - plane).max()))
