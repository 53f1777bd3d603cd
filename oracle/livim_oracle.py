"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

A restatement, on the same OpenCV (cv2) primitives the reference calls, of the per-frame hot path of
tschnz/Live-Video-Magnification behind ``MagnificationProcessor::process``:

  * src/processing/magnification/SpatialFilter.cpp   (pyramids, rolling window)
  * src/processing/magnification/TemporalFilter.cpp  (iirFilter, idealFilter, butterworth, Riesz IIR)
  * src/processing/magnification/RieszPyramid.cpp    (Riesz pyramid, phase diff, amplify, collapse)
  * src/processing/magnification/ComplexMat.hpp      (pair-of-Mat operator semantics)
  * src/processing/magnification/MagnifyCore.hpp     (magnifyMotion / magnifyColor / magnifyRiesz)
  * src/processing/MagnificationProcessor.cpp        (level clamp, structural reset, passthrough)
  * src/processing/MagnificationParamsUi.hpp         (UI Hz / % -> algorithm units)

Every function cites the reference file:line it follows. Only ``tests/`` (its helper scripts under ``tests/tools/``
included), ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module.

PARITY PIN STATUS: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c), and its CMake/vcpkg build cannot run here (no OpenCV C++ headers or libraries, no Qt).  The oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF RUN HERE instead: ``oracle/build_ref.py`` compiles the reference's
own hot-path sources where they lie (/root/reference/src/processing/**, unmodified) against ``oracle/cvshim`` —
an OpenCV-shaped facade that forwards every pixel operation to the real OpenCV 4.13 kernels in ``cv2`` — into
``oracle/_ref/_livim_ref``; ``tests/test_ref_pin.py`` requires this restatement to reproduce the compiled
reference BIT FOR BIT (u8 outputs, passthrough decisions, float temporal state of all three modes, the front
of the chain, and the host functions), and ``tests/golden/*.npz`` are outputs of that compiled reference
(generator: tests/golden/make_golden.py).  What remains outside the pin, because cv2 cannot execute it, is
stated in oracle/cvshim/opencv2/core.hpp: ``Mat::convertTo`` and the MatExpr-to-OpenCV-call lowering are
restated from OpenCV's documented behaviour in both the facade and this file; and the OpenCV *version* is the
wheel's 4.13.0, not whatever the reference's vcpkg baseline (``a5ac4c37…``, vcpkg.json:6) resolves to.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field, replace
from typing import List, Optional, Tuple

import cv2
import numpy as np

F32 = np.float32

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBM = None


def _libm():
    """glibc acosf / cosf / sinf as the reference calls them (oracle/libm_f32.c), built on first use with gcc
    into oracle/_ref/ (numpy's float32 ufuncs differ from libm in the last ulp on about a third of samples)."""
    global _LIBM
    if _LIBM is None:
        src, out = os.path.join(_HERE, "libm_f32.c"), os.path.join(_HERE, "_ref", "libm_f32.so")
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            try:
                os.makedirs(os.path.dirname(out), exist_ok=True)
                tmp = out + f".{os.getpid()}.tmp"
                subprocess.run(["gcc", "-O2", "-shared", "-fPIC", src, "-o", tmp, "-lm"], check=True)
                os.replace(tmp, out)
            except (OSError, subprocess.CalledProcessError):
                if not os.path.exists(out):   # a prebuilt (possibly older-stamped) library is still the same code
                    raise
        lib = ctypes.CDLL(out)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.livim_arccos_f32.argtypes = [fp, fp, ctypes.c_size_t]
        lib.livim_cossin_f32.argtypes = [fp, fp, fp, ctypes.c_size_t]
        lib.livim_arccos_f32.restype = lib.livim_cossin_f32.restype = None
        _LIBM = lib
    return _LIBM


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))

# --------------------------------------------------------------------------------------------------
# IProcessor.hpp:10-48 — parameter structs
# --------------------------------------------------------------------------------------------------
MODE_LAPLACE, MODE_PHASE, MODE_COLOR, MODE_NONE = 0, 1, 2, 3  # IProcessor.hpp:10


@dataclass
class MagnificationParams:  # IProcessor.hpp:14-23
    mode: int = MODE_LAPLACE
    amplification: float = 0.0
    coWavelength: float = 0.0
    coLow: float = 0.0
    coHigh: float = 0.0
    chromAttenuation: float = 0.0
    levels: int = 4
    framerate: float = 30.0


@dataclass
class PreprocessParams:  # IProcessor.hpp:26-41 (only takes part in the reset decision)
    downscale: int = 1
    roiEnabled: bool = False
    roiX: float = 0.0
    roiY: float = 0.0
    roiW: float = 1.0
    roiH: float = 1.0


@dataclass
class ProcessorConfig:  # IProcessor.hpp:44-48
    grayscale: bool = False
    preprocess: PreprocessParams = field(default_factory=PreprocessParams)
    magnification: MagnificationParams = field(default_factory=MagnificationParams)


# --------------------------------------------------------------------------------------------------
# MagnificationParamsUi.hpp — UI units -> algorithm units
# --------------------------------------------------------------------------------------------------
K_TWO_PI = 6.283185307179586  # MagnificationParamsUi.hpp:27


def motion_hz_to_blend(hz: float, fps: float) -> float:
    """MagnificationParamsUi.hpp:29-34."""
    if fps <= 0.0:
        fps = 30.0
    if hz <= 0.0:
        return 0.0
    a = 1.0 - math.exp(-K_TWO_PI * hz / fps)
    return min(max(a, 0.0), 0.999999)


def to_params(mode: int, amplification: int, wavelength: float, low: float, high: float,
              chroma: int, levels: int, fps: float) -> MagnificationParams:
    """MagnificationParamsUi.hpp:74-103 (toParams)."""
    p = MagnificationParams(mode=mode, amplification=float(amplification), levels=levels,
                            framerate=fps)
    if mode == MODE_COLOR:
        p.coWavelength, p.coLow, p.coHigh, p.chromAttenuation = 0.0, low, high, 0.0
    elif mode == MODE_LAPLACE:
        p.coWavelength = wavelength * 10.0
        p.coLow = motion_hz_to_blend(low, fps)
        p.coHigh = motion_hz_to_blend(high, fps)
        p.chromAttenuation = chroma / 100.0
    elif mode == MODE_PHASE:
        p.coWavelength = 100.0 - wavelength
        p.coLow, p.coHigh, p.chromAttenuation = low, high, 0.0
    return p


# --------------------------------------------------------------------------------------------------
# helpers that restate cv::Mat::convertTo (not reachable from Python)
# --------------------------------------------------------------------------------------------------
def _u8_to_f32_scaled(img8: np.ndarray, alpha: float) -> np.ndarray:
    """Mat::convertTo(CV_32F, alpha): OpenCV narrows alpha to float and computes x*a (+0)."""
    return img8.astype(F32) * F32(alpha)


def _f32_to_u8(img: np.ndarray, alpha: float, beta: float) -> np.ndarray:
    """Mat::convertTo(CV_8U, alpha, beta) on CV_32F: saturate(round_half_even(fma(x, (float)a, (float)b))).  The rounding
    is cvtps2dq: NaN and values outside the int32 range become INT_MIN, which then saturates to 0 (so +inf -> 0)."""
    a, b = float(F32(alpha)), float(F32(beta))
    with np.errstate(all="ignore"):
        v = (img.astype(np.float64) * a + b).astype(F32)  # one rounding == fmaf
        r = np.rint(v.astype(np.float64))
    bad = ~np.isfinite(r) | (r >= 2147483648.0) | (r < -2147483648.0)
    r = np.where(bad, -2147483648.0, r)
    return np.clip(r, 0, 255).astype(np.uint8)


def _scale(m: np.ndarray, s: float) -> np.ndarray:
    """MatExpr ``m * s`` on CV_32F lowers to convertTo(alpha=s): f32 multiply by (float)s."""
    return (m * F32(s)).astype(F32)


# --------------------------------------------------------------------------------------------------
# SpatialFilter.cpp
# --------------------------------------------------------------------------------------------------
def calculate_max_levels(w: int, h: int) -> int:
    """SpatialFilter.cpp:5-11."""
    if w > 5 and h > 5:
        return 1 + calculate_max_levels((1 + w) // 2, (1 + h) // 2)
    return 0


def build_gauss_pyr_from_img(img: np.ndarray, levels: int) -> List[np.ndarray]:
    """SpatialFilter.cpp:13-23."""
    pyr, cur = [], img
    for _ in range(levels):
        down = cv2.pyrDown(cur)
        pyr.append(down)
        cur = down
    return pyr


def build_laplace_pyr_from_img(img: np.ndarray, levels: int) -> List[np.ndarray]:
    """SpatialFilter.cpp:25-38."""
    pyr, cur = [], img
    for _ in range(levels):
        down = cv2.pyrDown(cur)
        up = cv2.pyrUp(down, dstsize=(cur.shape[1], cur.shape[0]))
        pyr.append(cv2.subtract(cur, up))
        cur = down
    pyr.append(cur)
    return pyr


def build_img_from_gauss_pyr(small: np.ndarray, levels: int, size_wh: Tuple[int, int]) -> np.ndarray:
    """SpatialFilter.cpp:40-50."""
    cur = small.copy()
    for _ in range(levels):
        cur = cv2.pyrUp(cur)
    return cv2.resize(cur, size_wh)


def build_img_from_laplace_pyr(pyr: List[np.ndarray], levels: int) -> np.ndarray:
    """SpatialFilter.cpp:52-61."""
    cur = pyr[levels]
    for level in range(levels - 1, -1, -1):
        up = cv2.pyrUp(cur, dstsize=(pyr[level].shape[1], pyr[level].shape[0]))
        cur = cv2.add(up, pyr[level])
    return cur.copy()


def img2temp_mat(frame: np.ndarray, dst: Optional[np.ndarray], max_images: int) -> np.ndarray:
    """SpatialFilter.cpp:63-84. Window layout here: (rows=h*w, cols=frames, C)."""
    c = 1 if frame.ndim == 2 else frame.shape[2]
    col = frame.reshape(-1, 1, c).astype(F32)
    dst = col.copy() if dst is None or dst.shape[1] == 0 else np.concatenate([dst, col], axis=1)
    if dst.shape[1] > max_images and max_images > 0:
        dst = dst[:, 1:].copy()
    return dst


def temp_mat2img(src: np.ndarray, position: int, h: int, w: int) -> np.ndarray:
    """SpatialFilter.cpp:86-89."""
    c = src.shape[2]
    out = src[:, position].reshape(h, w, c).copy()
    return out[..., 0] if c == 1 else out


# --------------------------------------------------------------------------------------------------
# TemporalFilter.cpp
# --------------------------------------------------------------------------------------------------
def iir_filter(src, lowpass_hi, lowpass_lo, cutoff_lo: float, cutoff_hi: float):
    """TemporalFilter.cpp:9-22. Returns (dst, new_hi, new_lo)."""
    if cutoff_lo == 0:
        cutoff_lo = 0.01
    tmp1 = cv2.addWeighted(lowpass_hi, 1 - cutoff_hi, src, cutoff_hi, 0)
    tmp2 = cv2.addWeighted(lowpass_lo, 1 - cutoff_lo, src, cutoff_lo, 0)
    return cv2.subtract(tmp1, tmp2), tmp1, tmp2


def get_optimal_buffer_size(fps: int) -> int:
    """TemporalFilter.cpp:82-94."""
    r = max(2 * fps, 16) & 0xFFFFFFFF
    r -= 1
    for s in (1, 2, 4, 8, 16):
        r |= r >> s
    return (r + 1) & 0xFFFFFFFF


def create_ideal_bandpass_filter(rows: int, width: int, cutoff_lo: float, cutoff_hi: float,
                                 framerate: float) -> np.ndarray:
    """TemporalFilter.cpp:59-80: real 0/1 mask over *packed* indices."""
    wf = float(F32(width))
    fl = 2 * cutoff_lo * wf / framerate
    fh = 2 * cutoff_hi * wf / framerate
    x = np.arange(width)
    row = ((x >= fl) & (x <= fh)).astype(F32)
    return np.tile(row, (rows, 1))


def ideal_filter(src: np.ndarray, cutoff_lo: float, cutoff_hi: float, framerate: float) -> np.ndarray:
    """TemporalFilter.cpp:24-57. src: (rows, cols, C) f32."""
    if cutoff_lo == 0.0:
        cutoff_lo += 0.01
    chans = []
    for c in range(src.shape[2]):
        current = np.ascontiguousarray(src[:, :, c])
        height = cv2.getOptimalDFTSize(current.shape[0])
        temp = cv2.copyMakeBorder(current, 0, height - current.shape[0], 0, 0,
                                  cv2.BORDER_CONSTANT, value=0)
        temp = cv2.dft(temp, flags=cv2.DFT_ROWS | cv2.DFT_SCALE)
        filt = create_ideal_bandpass_filter(temp.shape[0], temp.shape[1], cutoff_lo, cutoff_hi,
                                            framerate)
        temp = cv2.mulSpectrums(temp, filt, cv2.DFT_ROWS)
        temp = cv2.idft(temp, flags=cv2.DFT_ROWS | cv2.DFT_SCALE)
        chans.append(temp[:current.shape[0], :current.shape[1]].copy())
    dst = np.stack(chans, axis=2)
    # cv::normalize(dst, dst, 0, 1, NORM_MINMAX): global over all channels
    flat = np.ascontiguousarray(dst).reshape(dst.shape[0], -1)
    flat = cv2.normalize(flat, None, 0, 1, cv2.NORM_MINMAX)
    return flat.reshape(dst.shape)


def butterworth(n_order: int, wn: float) -> Tuple[List[float], List[float]]:
    """TemporalFilter.cpp:96-297 (butterworth + helpers), restated step by step in complex128.

    Returns (a, b).  For N=2 equals scipy.signal.butter(2, Wn) to ~1e-16 (tests check this).
    """
    fs = 2.0
    w0 = 2.0 * fs * math.tan(math.pi * wn / fs)
    # prototypeAnalogButterworth :268-277
    j = 1j
    poles = [np.exp(j * (2.0 * k - 1) / (2.0 * n_order) * math.pi) * j for k in range(1, n_order + 1)]
    zeros: List[complex] = []
    gain = 1.0

    def sort_key(z):  # sortComplex :99-105
        return (z.real, z.imag)

    def poly_coeffs(roots):  # polynomialCoefficients :110-146
        roots = sorted(roots, key=sort_key)
        coeffs = [0j] * (len(roots) + 1)
        coeffs[0] = 1.0 + 0j
        sofar = 1
        for r in roots:
            w = -r
            for jx in range(sofar, 0, -1):
                coeffs[jx] = coeffs[jx] * w + coeffs[jx - 1]
            coeffs[0] *= w
            sofar += 1
        pos = sorted([r for r in roots if not r.imag < 0], key=sort_key)
        neg = sorted([r for r in roots if not r.imag > 0], key=sort_key)
        result = list(coeffs)
        if len(pos) == len(neg) and all(p == q for p, q in zip(pos, neg)):
            result = [complex(c.real, 0.0) for c in coeffs]
        return result

    a = poly_coeffs(poles)
    b = [c * gain for c in poly_coeffs(zeros)]

    def normalize(bb, aa):  # :160-167
        lead = aa[0]
        aa[:] = [0j if lead == 0 else x / lead for x in aa]
        bb[:] = [0j if lead == 0 else x / lead for x in bb]

    # toLowpass :234-264
    d, n = len(a), len(b)
    m = max(d, n)
    start1, start2 = max(n - d, 0), max(d - n, 0)
    pwo = [w0 ** float(k) for k in range(m - 1, -1, -1)]
    k = start2
    while k < len(pwo) and k - start2 < len(b):
        if pwo[k] == 0.0:
            b[k - start2] = 0j
        else:
            b[k - start2] *= complex(pwo[start1]) / complex(pwo[k])
        k += 1
    k = start1
    while k < len(pwo) and k - start1 < len(a):
        if pwo[k] == 0.0:
            a[k - start1] = 0j
        else:
            a[k - start1] *= complex(pwo[start1]) / complex(pwo[k])
        k += 1
    normalize(b, a)

    # bilinearTransform :187-230
    def choose(nn, kk):  # :170-183
        return math.comb(nn, kk) if kk <= nn else 0

    dd, nn = len(a) - 1, len(b) - 1
    mm = max(nn, dd)

    def transform(c, deg):
        out = []
        for jx in range(mm + 1):
            val = 0j
            for i in range(deg + 1):
                for kx in range(i + 1):
                    for lx in range(mm - i + 1):
                        if kx + lx == jx:
                            val += (complex(choose(i, kx)) * complex(choose(mm - i, lx)) * c[deg - i]
                                    * (2.0 * fs) ** i * (-1.0) ** kx)
            out.append(complex(val.real, 0.0))
        return out

    bprime, aprime = transform(b, nn), transform(a, dd)
    normalize(bprime, aprime)
    return [x.real for x in aprime], [x.real for x in bprime]


def _mul_scalar(m: np.ndarray, s: float) -> np.ndarray:
    """cv::multiply(Mat, double) used by ComplexMat.hpp:47-52: f32(double(x) * s)."""
    return cv2.multiply(m, float(s))


class RieszTemporalFilter:
    """TemporalFilter.cpp:299-362. State per level: phase, register0, register1, each a (cos, sin) pair."""

    def __init__(self, frq: float, fps: float, sizes_hw: List[Tuple[int, int]]):
        self.frequency, self.framerate = frq, fps
        self.A: List[float] = []
        self.B: List[float] = []
        z = lambda hw: [np.zeros(hw, F32), np.zeros(hw, F32)]
        self.reg0 = [z(s) for s in sizes_hw]
        self.reg1 = [z(s) for s in sizes_hw]
        self.phase = [z(s) for s in sizes_hw]

    def compute_coefficients(self):  # :324-327
        wn = 0.0 if self.framerate == 0.0 else self.frequency / (self.framerate / 2.0)
        self.A, self.B = butterworth(2, wn)

    def update_frequency(self, f: float):  # :319-322
        self.frequency = f
        self.compute_coefficients()

    def reset_mat(self):  # :353-362
        for grp in (self.reg0, self.reg1, self.phase):
            for pair in grp:
                pair[0][:] = 0
                pair[1][:] = 0

    def iir_temporal_filter(self, phase_diff, lvl: int):
        """:340-351 (Direct Form II). Returns result pair."""
        res = [None, None]
        for k in range(2):
            self.phase[lvl][k] = cv2.add(self.phase[lvl][k], phase_diff[k])
            ph = self.phase[lvl][k]
            res[k] = cv2.add(_mul_scalar(ph, self.B[0]), self.reg0[lvl][k])
            self.reg0[lvl][k] = cv2.subtract(cv2.add(_mul_scalar(ph, self.B[1]), self.reg1[lvl][k]),
                                             _mul_scalar(res[k], self.A[1]))
            self.reg1[lvl][k] = cv2.subtract(_mul_scalar(ph, self.B[2]), _mul_scalar(res[k], self.A[2]))
        return res


# --------------------------------------------------------------------------------------------------
# RieszPyramid.cpp
# --------------------------------------------------------------------------------------------------
# 9x9 tap tables, RieszPyramid.cpp:146-167 (literal 4-decimal constants of the Riesz-pyramid paper).
_LP_Q = [  # unique quadrant rows (0..4) x cols (0..4); table is symmetric under flips
    [-0.0001, -0.0007, -0.0023, -0.0046, -0.0057],
    [-0.0007, -0.0030, -0.0047, -0.0025, -0.0003],
    [-0.0023, -0.0047, 0.0054, 0.0272, 0.0387],
    [-0.0046, -0.0025, 0.0272, 0.0706, 0.0910],
    [-0.0057, -0.0003, 0.0387, 0.0910, 0.1138],
]
_HP_Q = [
    [0.0000, 0.0003, 0.0011, 0.0022, 0.0027],
    [0.0003, 0.0020, 0.0059, 0.0103, 0.0123],
    [0.0011, 0.0059, 0.0151, 0.0249, 0.0292],
    [0.0022, 0.0103, 0.0249, 0.0402, 0.0469],
    [0.0027, 0.0123, 0.0292, 0.0469, -0.9455],
]


def _full9(q):
    q = np.array(q, dtype=F32)
    top = np.concatenate([q, q[:, 3::-1]], axis=1)
    return np.concatenate([top, top[3::-1]], axis=0)


LOWPASS_9x9 = _full9(_LP_Q)
HIGHPASS_9x9 = _full9(_HP_Q)
RIESZ_TAPS = np.array([[-0.2, -0.48, 0, 0.48, 0.2]], dtype=F32)  # RieszPyramid.cpp:71


def _filter2d(img, k):
    return cv2.filter2D(img, cv2.CV_32F, k, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REFLECT_101)


def arc_cos(x: np.ndarray) -> np.ndarray:
    """RieszPyramid.cpp:8-23 (libm acosf) — note the clamp returns -1.0/+1.0 *radians* (quirk)."""
    x = np.ascontiguousarray(x, dtype=F32)
    out = np.empty_like(x)
    _libm().livim_arccos_f32(_fptr(x), _fptr(out), x.size)
    return out


def cos_sin(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """RieszPyramid.cpp:25-38 (libm cosf / sinf)."""
    x = np.ascontiguousarray(x, dtype=F32)
    c, s_ = np.empty_like(x), np.empty_like(x)
    _libm().livim_cossin_f32(_fptr(x), _fptr(c), _fptr(s_), x.size)
    return c, s_


def _patch_nans(m: np.ndarray) -> np.ndarray:
    m = m.copy()
    m[np.isnan(m)] = 0.0  # cv::patchNaNs: NaN only, inf untouched
    return m


class RieszLevel:
    """RieszPyramidLevel (RieszPyramid.hpp:19-55)."""

    def __init__(self):
        self.lowpass = None
        self.rx = None  # real(itsRiesz)  (1x5 filter)
        self.ry = None  # imag(itsRiesz)  (5x1 filter)
        self.amplitude = None
        self.amplitude_blurred = None
        self.phase_diff = [None, None]
        self.highpass_iir = [None, None]
        self.lowpass_iir = [None, None]

    def copy_from(self, o: "RieszLevel"):
        """operator= (RieszPyramid.cpp:52-64): lowpass, riesz, phaseDiff, amplitude(s) are deep-copied."""
        cp = lambda m: None if m is None else m.copy()
        self.lowpass, self.rx, self.ry = cp(o.lowpass), cp(o.rx), cp(o.ry)
        self.phase_diff = [cp(o.phase_diff[0]), cp(o.phase_diff[1])]
        self.amplitude, self.amplitude_blurred = cp(o.amplitude), cp(o.amplitude_blurred)

    def build(self, octave: np.ndarray):
        """RieszPyramid.cpp:66-78."""
        self.lowpass = octave
        self.rx = _filter2d(octave, RIESZ_TAPS)
        self.ry = _filter2d(octave, RIESZ_TAPS.T.copy())

    def compute_phase_difference_and_amplitude(self, prior: "RieszLevel"):
        """RieszPyramid.cpp:81-111."""
        mul, add = cv2.multiply, cv2.add
        q_real = add(add(mul(self.lowpass, prior.lowpass), mul(self.rx, prior.rx)),
                     mul(self.ry, prior.ry))
        neg_low = _scale(self.lowpass, -1.0)
        qx = add(mul(prior.rx, neg_low), mul(self.rx, prior.lowpass))
        qy = add(mul(prior.ry, neg_low), mul(self.ry, prior.lowpass))
        xy_sq = add(mul(qx, qx), mul(qy, qy))
        with np.errstate(all="ignore"):
            q_amp = cv2.sqrt(add(mul(q_real, q_real), xy_sq))
            tmp = cv2.divide(q_real, q_amp)
            phase_difference = arc_cos(tmp)
            xy_sqrt = cv2.sqrt(xy_sq)
            ox, oy = cv2.divide(qx, xy_sqrt), cv2.divide(qy, xy_sqrt)
            self.phase_diff = [_patch_nans(mul(ox, phase_difference)),
                               _patch_nans(mul(oy, phase_difference))]
            self.amplitude = cv2.sqrt(q_amp)
        self.amplitude_blurred = cv2.GaussianBlur(self.amplitude, (13, 13), 3.0)

    def normalize(self):
        """RieszPyramid.cpp:114-127."""
        kernel = cv2.getGaussianKernel(13, 3.0, cv2.CV_32F)
        out = []
        for k in range(2):
            change = cv2.subtract(self.highpass_iir[k], self.lowpass_iir[k])
            r = cv2.multiply(change, self.amplitude)
            r = cv2.sepFilter2D(r, -1, kernel, kernel, anchor=(-1, -1), delta=0,
                                borderType=cv2.BORDER_REFLECT_101)
            with np.errstate(all="ignore"):
                r = cv2.divide(r, self.amplitude_blurred)
            out.append(r)
        return out

    def amplify(self, alpha: float, threshold: float):
        """RieszPyramid.cpp:129-144."""
        tc, ts = self.normalize()
        with np.errstate(all="ignore"):
            mag_v = cv2.sqrt(cv2.add(cv2.multiply(tc, tc), cv2.multiply(ts, ts)))
            mag_v2 = _scale(mag_v, alpha)
            _, mag_v2 = cv2.threshold(mag_v2, threshold, 0, cv2.THRESH_TRUNC)
            pc, ps = cos_sin(mag_v2)
            pair = cv2.add(cv2.multiply(self.rx, tc), cv2.multiply(self.ry, ts))
            pair = _patch_nans(cv2.divide(pair, mag_v))
            self.lowpass = cv2.subtract(cv2.multiply(self.lowpass, pc), cv2.multiply(pair, ps))


class RieszPyramid:
    """RieszPyramid (RieszPyramid.hpp:57-91)."""

    def __init__(self):
        self.num_levels = 0
        self.levels: List[RieszLevel] = []

    def init(self, frame: np.ndarray, levels: int):
        """RieszPyramid.cpp:196-213 — NB: zeroes the Riesz pair after buildPyramid (quirk)."""
        self.levels = [RieszLevel() for _ in range(levels)]
        self.num_levels = levels
        self.build_pyramid(frame)
        for lv in self.levels:
            z = lambda: np.zeros(lv.lowpass.shape, F32)
            lv.rx, lv.ry = z(), z()
            lv.phase_diff = [z(), z()]
            lv.lowpass_iir = [z(), z()]
            lv.highpass_iir = [z(), z()]
            lv.amplitude, lv.amplitude_blurred = z(), z()

    @staticmethod
    def subsample(img: np.ndarray) -> np.ndarray:
        """RieszPyramid.cpp:254-278: keep even rows/cols."""
        return img[::2, ::2].copy()

    @staticmethod
    def inject_zeros_even(img: np.ndarray) -> np.ndarray:
        """RieszPyramid.cpp:280-302."""
        out = np.zeros_like(img)
        out[::2, ::2] = img[::2, ::2]
        return out

    def build_pyramid(self, frame: np.ndarray):
        """RieszPyramid.cpp:215-238."""
        mx = self.num_levels - 1
        if mx == -1:
            return
        octave = frame
        for i in range(mx):
            hp = _filter2d(octave, HIGHPASS_9x9)
            self.levels[i].build(hp)
            lp = _filter2d(octave, (LOWPASS_9x9 * F32(2.0)).astype(F32))
            octave = self.subsample(lp)
        self.levels[mx].build(octave)

    def sizes(self):
        return [lv.lowpass.shape for lv in self.levels]

    def compute_phase_difference_and_amplitude(self, prior: "RieszPyramid"):
        """RieszPyramid.cpp:240-246."""
        for i in range(len(self.levels) - 1):
            self.levels[i].compute_phase_difference_and_amplitude(prior.levels[i])

    def amplify(self, alpha: float, threshold: float):
        """RieszPyramid.cpp:248-252."""
        for i in range(self.num_levels - 2, -1, -1):
            self.levels[i].amplify(alpha, threshold)

    def collapse_pyramid(self) -> np.ndarray:
        """RieszPyramid.cpp:304-325."""
        count = len(self.levels) - 1
        result = self.levels[count].lowpass
        for i in range(count - 1, -1, -1):
            octave = self.levels[i].lowpass
            up = cv2.resize(result, (octave.shape[1], octave.shape[0]), interpolation=cv2.INTER_NEAREST)
            up_zero = self.inject_zeros_even(up)
            lp = _filter2d(up_zero, (LOWPASS_9x9 * F32(2.0)).astype(F32))
            hp = _filter2d(octave, HIGHPASS_9x9)
            result = cv2.add(lp, hp)
        return result

    def assign(self, other: "RieszPyramid"):
        """operator= (RieszPyramid.cpp:182-194)."""
        self.num_levels = other.num_levels
        if len(self.levels) != len(other.levels):
            self.levels = [RieszLevel() for _ in other.levels]
        for a, b in zip(self.levels, other.levels):
            a.copy_from(b)


# --------------------------------------------------------------------------------------------------
# MagnifyCore.hpp — state + per-frame drivers
# --------------------------------------------------------------------------------------------------
class MotionState:  # MagnifyCore.hpp:24-29
    def __init__(self):
        self.lowpassHi: List[np.ndarray] = []
        self.lowpassLo: List[np.ndarray] = []

    def empty(self):
        return len(self.lowpassHi) == 0

    def reset(self):
        self.lowpassHi, self.lowpassLo = [], []


class ColorState:  # MagnifyCore.hpp:31-34
    def __init__(self):
        self.window: Optional[np.ndarray] = None

    def reset(self):
        self.window = None


class RieszState:  # MagnifyCore.hpp:36-40
    def __init__(self):
        self.cur = self.old = self.lo = self.hi = None

    def reset(self):
        self.cur = self.old = self.lo = self.hi = None


def motion_gains(p: MagnificationParams, levels: int, w: int, h: int) -> List[float]:
    """MagnifyCore.hpp:114-134: per-level gain (index = pyramid level; float arithmetic as written)."""
    delta = F32(p.coWavelength / (8.0 * (1.0 + p.amplification)))
    exaggeration = F32(2.0)
    lam = F32(math.sqrt(float(w * w + h * h)) / 3.0)
    gains = [0.0] * (levels + 1)
    for cur in range(levels, -1, -1):
        with np.errstate(all="ignore"):
            # C++ double arithmetic: x/0 = inf (coWavelength == 0), no exception
            curr_alpha = F32((np.float64(lam) / (np.float64(delta) * 8.0) - 1.0) * np.float64(exaggeration))
        gains[cur] = 0.0 if cur in (levels, 0) else float(min(F32(p.amplification), curr_alpha))
        lam = F32(float(lam) / 2.0)
    return gains


def magnify_motion(in8u: np.ndarray, p: MagnificationParams, levels: int, channels: int,
                   st: MotionState, debug: Optional[dict] = None):
    """MagnifyCore.hpp:83-160. Returns (produced, out8u)."""
    color = channels >= 3
    inp = _u8_to_f32_scaled(in8u, 1.0 / 255.0)  # :89 / :92
    if color:
        inp = cv2.cvtColor(inp, cv2.COLOR_BGR2Lab)  # :90
    pyramid = build_laplace_pyr_from_img(inp, levels)  # :96
    if st.empty():  # :98-103
        st.lowpassHi = [m.copy() for m in pyramid]
        st.lowpassLo = [m.copy() for m in pyramid]
        output = inp
    else:
        motion = [None] * (levels + 1)
        for lv in range(levels):  # :106-109
            motion[lv], st.lowpassHi[lv], st.lowpassLo[lv] = iir_filter(
                pyramid[lv], st.lowpassHi[lv], st.lowpassLo[lv], p.coLow, p.coHigh)
        motion[levels] = pyramid[levels]  # :112
        h, w = inp.shape[:2]
        gains = motion_gains(p, levels, w, h)
        for lv in range(levels, -1, -1):  # :127-134
            motion[lv] = _scale(motion[lv], gains[lv])
        mot = build_img_from_laplace_pyr(motion, levels)  # :136-137
        if color:  # :140-146
            planes = list(cv2.split(mot))
            planes[1] = _scale(planes[1], p.chromAttenuation)
            planes[2] = _scale(planes[2], p.chromAttenuation)
            mot = cv2.merge(planes)
        output = cv2.add(inp, mot)  # :148
        if debug is not None:
            debug["motion"] = mot
    if debug is not None:
        debug["input_f32"], debug["output_f32"], debug["pyramid"] = inp, output, pyramid
    if color:  # :151-153
        output = cv2.cvtColor(output, cv2.COLOR_Lab2BGR)
    if debug is not None:
        debug["output_bgr_f32"] = output
    return True, _f32_to_u8(output, 255.0, 1.0 / 255.0)  # :153 / :156


def magnify_color(in8u: np.ndarray, p: MagnificationParams, levels: int, channels: int,
                  st: ColorState, debug: Optional[dict] = None):
    """MagnifyCore.hpp:163-206."""
    inp = in8u.astype(F32)  # :168-169 (no 1/255, no Lab)
    pyr = build_gauss_pyr_from_img(inp, levels)  # :171-172
    small = pyr[levels - 1]
    st.window = img2temp_mat(small, st.window, get_optimal_buffer_size(int(p.framerate)))  # :176
    if st.window.shape[1] < 2:  # :180
        return False, None
    filtered = ideal_filter(st.window, p.coLow, p.coHigh, p.framerate)  # :183
    filtered = _scale(filtered, p.amplification)  # :185
    sh, sw = small.shape[:2]
    frame = temp_mat2img(filtered, min(1, filtered.shape[1] - 1), sh, sw)  # :189-192
    color_img = build_img_from_gauss_pyr(frame, levels, (inp.shape[1], inp.shape[0]))  # :194-195
    output = cv2.add(inp, color_img)  # :197
    mn, mx = float(output.min()), float(output.max())  # :200-201
    if debug is not None:
        debug["output_f32"], debug["min"], debug["max"] = output, mn, mx
        debug["filtered"] = filtered
    with np.errstate(all="ignore"):
        # C++ double arithmetic: a constant image gives 255/0 = inf and -min*255/0 = -inf or NaN, no exception
        span = np.float64(mx) - np.float64(mn)
        out8 = _f32_to_u8(output, float(np.float64(255.0) / span), float(np.float64(-mn) * 255.0 / span))  # :202-203
    return True, out8


PI_PERCENT = math.pi / 100.0  # MagnifyCore.hpp:214


def magnify_riesz(in8u: np.ndarray, p: MagnificationParams, levels: int, channels: int,
                  st: RieszState, debug: Optional[dict] = None):
    """MagnifyCore.hpp:209-279."""
    if channels < 3:  # :212
        return False, None
    buf = _u8_to_f32_scaled(in8u, 1.0 / 255.0)  # :218
    buf = cv2.cvtColor(buf, cv2.COLOR_BGR2Lab)  # :219
    lab = list(cv2.split(buf))
    inp = lab[0]
    if st.cur is None or math.isnan(st.lo.A[0]) or math.isnan(st.hi.A[0]):  # :226-240
        st.reset()
        st.cur, st.old = RieszPyramid(), RieszPyramid()
        st.cur.init(inp, levels)
        st.old.init(inp, levels)
        st.lo = RieszTemporalFilter(p.coLow, p.framerate, st.cur.sizes())
        st.hi = RieszTemporalFilter(p.coHigh, p.framerate, st.cur.sizes())
        st.lo.compute_coefficients()
        st.hi.compute_coefficients()
        return False, None
    if st.lo.frequency != p.coLow:  # :243-248
        st.lo.update_frequency(p.coLow)
        st.lo.reset_mat()
        st.hi.reset_mat()
        st.old.build_pyramid(inp)
    if st.hi.frequency != p.coHigh:  # :249-254
        st.hi.update_frequency(p.coHigh)
        st.hi.reset_mat()
        st.lo.reset_mat()
        st.old.build_pyramid(inp)
    st.cur.build_pyramid(inp)  # :256
    st.cur.compute_phase_difference_and_amplitude(st.old)  # :257
    for lvl in range(st.cur.num_levels - 1):  # :259-264
        st.cur.levels[lvl].lowpass_iir = st.lo.iir_temporal_filter(st.cur.levels[lvl].phase_diff, lvl)
        st.cur.levels[lvl].highpass_iir = st.hi.iir_temporal_filter(st.cur.levels[lvl].phase_diff, lvl)
    st.old.assign(st.cur)  # :267
    if debug is not None:
        debug["band_lowpass"] = [lv.lowpass.copy() for lv in st.cur.levels]
        debug["amplitude"] = [None if lv.amplitude is None else lv.amplitude.copy() for lv in st.cur.levels]
        debug["phase_diff"] = [[None if m is None else m.copy() for m in lv.phase_diff] for lv in st.cur.levels]
    st.cur.amplify(p.amplification, p.coWavelength * PI_PERCENT)  # :269
    magnified = st.cur.collapse_pyramid()  # :270
    if debug is not None:
        debug["magnified_L"] = magnified.copy()
        debug["amplified_lowpass"] = [lv.lowpass.copy() for lv in st.cur.levels]
    lab[0] = magnified.astype(F32)
    output = cv2.cvtColor(cv2.merge(lab), cv2.COLOR_Lab2BGR)  # :272-275
    if debug is not None:
        debug["output_bgr_f32"] = output
    return True, _f32_to_u8(output, 255.0, 1.0 / 255.0)  # :276


# --------------------------------------------------------------------------------------------------
# MagnifyCore.hpp:45-80 StructuralTracker + MagnificationProcessor.cpp
# --------------------------------------------------------------------------------------------------
class StructuralTracker:
    def __init__(self):
        self.reset()

    def update(self, cfg: ProcessorConfig, lv: int, ch: int, size_wh) -> bool:  # :53-65
        p = cfg.magnification
        change = (p.mode != self.mode or lv != self.levels or size_wh != self.size
                  or ch != self.channels or cfg.preprocess != self.preprocess)
        if change:
            self.mode, self.levels, self.size, self.channels = p.mode, lv, size_wh, ch
            self.preprocess = replace(cfg.preprocess)
        return change

    def disable(self):  # :68-73
        self.mode, self.levels, self.channels, self.size = MODE_NONE, -1, -1, (0, 0)

    def reset(self):  # :76-79
        self.disable()
        self.preprocess = PreprocessParams()


class MagnificationProcessor:
    """MagnificationProcessor.cpp:10-67. ``process`` returns (produced, image): produced=False means the
    reference returns the *input* FrameRef unchanged (identity / passthrough)."""

    def __init__(self):
        self.tracker = StructuralTracker()
        self.motion, self.color, self.riesz = MotionState(), ColorState(), RieszState()

    def reset(self):  # :10-15
        self.motion.reset(); self.color.reset(); self.riesz.reset(); self.tracker.reset()

    def process(self, image: np.ndarray, cfg: ProcessorConfig, debug: Optional[dict] = None):
        p = cfg.magnification
        if p.mode == MODE_NONE or image is None or image.size == 0:  # :21-29
            if self.tracker.mode != MODE_NONE:
                self.motion.reset(); self.color.reset(); self.riesz.reset(); self.tracker.disable()
            return False, image
        h, w = image.shape[:2]
        max_levels = calculate_max_levels(w, h)  # :32
        if max_levels < 1:
            return False, image
        levels = min(max(p.levels, 1), max_levels)  # :34
        channels = 1 if image.ndim == 2 else image.shape[2]
        if self.tracker.update(cfg, levels, channels, (w, h)):  # :39-43
            self.motion.reset(); self.color.reset(); self.riesz.reset()
        if p.mode == MODE_LAPLACE:
            produced, out = magnify_motion(image, p, levels, channels, self.motion, debug)
        elif p.mode == MODE_COLOR:
            produced, out = magnify_color(image, p, levels, channels, self.color, debug)
        else:
            produced, out = magnify_riesz(image, p, levels, channels, self.riesz, debug)
        if not produced:  # :61
            return False, image
        return True, out


# --------------------------------------------------------------------------------------------------
# Front of the chain: PreprocessProcessor.cpp, GrayscaleProcessor.cpp, ChainBuilder.cpp
# --------------------------------------------------------------------------------------------------
def _lround(x: float) -> int:
    """std::lround: half away from zero."""
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def preprocess(image: np.ndarray, cfg: ProcessorConfig):
    """PreprocessProcessor::process (PreprocessProcessor.cpp:10-51). Returns (changed, image)."""
    if image is None or image.size == 0:
        return False, image
    p = cfg.preprocess
    divisor = min(max(p.downscale, 1), 8)
    if not p.roiEnabled and divisor == 1:
        return False, image
    rows, cols = image.shape[:2]
    x, y, w, h = 0, 0, cols, rows
    if p.roiEnabled:
        f = lambda v: float(np.float32(v))   # the reference stores the ROI as float
        x, y = _lround(f(p.roiX) * cols), _lround(f(p.roiY) * rows)
        w, h = _lround(f(p.roiW) * cols), _lround(f(p.roiH) * rows)
        x = min(max(x, 0), cols - 1); y = min(max(y, 0), rows - 1)
        w = min(max(w, 1), cols - x); h = min(max(h, 1), rows - y)
    cropped = image[y:y + h, x:x + w]
    if divisor > 1:
        dw, dh = max(1, cropped.shape[1] // divisor), max(1, cropped.shape[0] // divisor)
        out = cv2.resize(np.ascontiguousarray(cropped), (dw, dh), interpolation=cv2.INTER_AREA)
    else:
        out = cropped.copy()
    return True, out


def grayscale(image: np.ndarray, cfg: ProcessorConfig):
    """GrayscaleProcessor::process (GrayscaleProcessor.cpp:7-16). Returns (changed, image)."""
    if not cfg.grayscale or image is None or image.size == 0 or image.ndim == 2:
        return False, image
    return True, cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)


def run_chain_once(magnifier: "MagnificationProcessor", image: np.ndarray, cfg: ProcessorConfig):
    """runChainOnce (ChainBuilder.cpp:19-29): Preprocess -> Grayscale -> Magnification; returns
    (cur, original, cur_is_input, original_is_input)."""
    pre_changed, cur = preprocess(image, cfg)
    original = cur
    gray_changed, cur = grayscale(cur, cfg)
    produced, out = magnifier.process(cur, cfg)
    if produced:
        cur = out
    return cur, original, not (pre_changed or gray_changed or produced), not pre_changed
