"""numpy statement of the exact arithmetic the CUDA kernels implement (operation order included).
Used by the CPU tests to check those formulas against cv2 before any GPU is involved, and by GPU
tests as a second, cv2-independent comparison.  Test infrastructure only."""
import numpy as np

F32 = np.float32


def reflect101(i, n):
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    i = np.where(i >= n, 2 * n - 2 - i, i)
    return np.clip(i, 0, n - 1)


def pyr_down(img):
    """cv::pyrDown, SURVEY A.1: horizontal [1 4 6 4 1] then vertical, scaled by 1/256."""
    img = img.astype(F32)
    h, w = img.shape[:2]
    ho, wo = (h + 1) // 2, (w + 1) // 2
    xs = [reflect101(2 * np.arange(wo) + d, w) for d in (-2, -1, 0, 1, 2)]
    row = img[:, xs[2]] * F32(6) + (img[:, xs[1]] + img[:, xs[3]]) * F32(4) + img[:, xs[0]] + img[:, xs[4]]
    ys = [reflect101(2 * np.arange(ho) + d, h) for d in (-2, -1, 0, 1, 2)]
    out = (row[ys[2]] * F32(6) + (row[ys[1]] + row[ys[3]]) * F32(4) + row[ys[0]] + row[ys[4]]) * F32(1 / 256)
    return out.astype(F32)


def _up_idx(n_f, n_c):
    """For fine index x: even -> (s[i-1], s[i], s[i+1]) ; odd -> (s[i], s[i+1]) with
    s[-1]:=s[1], s[n]:=s[n-1]."""
    x = np.arange(n_f)
    i = x // 2
    im1 = np.where(i - 1 < 0, 1, i - 1)
    ip1 = np.where(i + 1 >= n_c, n_c - 1, i + 1)
    return x % 2 == 1, im1, i, ip1


def pyr_up(img, dst_hw):
    img = img.astype(F32)
    hc, wc = img.shape[:2]
    hf, wf = dst_hw
    odd, im1, i0, ip1 = _up_idx(wf, wc)
    even_v = img[:, im1] + img[:, i0] * F32(6) + img[:, ip1]
    odd_v = (img[:, i0] + img[:, ip1]) * F32(4)
    sel = odd if img.ndim == 2 else odd[:, None]
    row = np.where(sel, odd_v, even_v)
    odd, im1, i0, ip1 = _up_idx(hf, hc)
    even_v = (row[im1] + row[i0] * F32(6) + row[ip1]) * F32(1 / 64)
    odd_v = ((row[i0] + row[ip1]) * F32(4)) * F32(1 / 64)
    sel = odd[:, None] if img.ndim == 2 else odd[:, None, None]
    return np.where(sel, odd_v, even_v).astype(F32)


def iir(x, hi, lo, c_lo, c_hi):
    """f32 result of an f64 weighted sum (cv::addWeighted)."""
    if c_lo == 0:
        c_lo = 0.01
    nh = (hi.astype(np.float64) * (1 - c_hi) + x.astype(np.float64) * c_hi).astype(F32)
    nl = (lo.astype(np.float64) * (1 - c_lo) + x.astype(np.float64) * c_lo).astype(F32)
    return (nh - nl).astype(F32), nh, nl
