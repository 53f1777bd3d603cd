// TEST INFRASTRUCTURE ONLY.  Compiles the product's __host__ __device__ per-pixel functions
// (csrc/mc_math.cuh) and host tables (csrc/mc_tables.cpp) for the CPU so the non-GPU test-suite can
// compare them with the oracle.  Never loaded by the product; the product has no CPU path.
#include <cmath>
#include <cstdint>
#include <vector>

#include "mc_internal.h"

using namespace mc;

namespace {
struct Tables {
    std::vector<LabLutCell> lut;
    std::vector<float4> gam;
    LabInvCoeffs k;
    Tables() {
        build_lab_lut_cells(lut);
        build_inv_gamma_spline(gam);
        build_lab_inv_coeffs(k);
    }
};
const Tables& T() {
    static Tables t;
    return t;
}
}  // namespace

extern "C" {
int hc_lut_entries() { return (int)T().lut.size(); }
void hc_bgr_to_lab(const uint8_t* bgr, int n, float* lab) {
    for (int i = 0; i < n; ++i)
        bgr_u8_to_lab(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2], T().lut.data(), lab[3 * i], lab[3 * i + 1], lab[3 * i + 2]);
}
void hc_lab_to_bgr(const float* lab, int n, float* bgr) {
    for (int i = 0; i < n; ++i)
        lab_to_bgr(lab[3 * i], lab[3 * i + 1], lab[3 * i + 2], T().k, T().gam.data(), bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
}
void hc_unit_to_u8(const float* x, int n, uint8_t* out) {
    for (int i = 0; i < n; ++i) out[i] = unit_to_u8(x[i]);
}
void hc_ema(const float* state, const float* x, int n, double c, float* out) {
    for (int i = 0; i < n; ++i) out[i] = ema(state[i], x[i], 1 - c, c);
}
void hc_gauss13(float* taps) { gaussian_kernel_13_3(taps); }
void hc_bgr2gray(const uint8_t* bgr, int n, uint8_t* gray) {
    for (int i = 0; i < n; ++i) gray[i] = bgr_to_gray_u8(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
}
// cv::resize(src, (dw, dh), INTER_AREA) through the product's per-sample function and tap builder
void hc_resize_area(const uint8_t* src, int h, int w, int cn, int dw, int dh, uint8_t* dst) {
    const double sx = (double)w / dw, sy = (double)h / dh;
    const int isx = (int)sx, isy = (int)sy;
    const bool fast = std::abs(sx - isx) < 2.220446049250313e-16 && std::abs(sy - isy) < 2.220446049250313e-16;
    std::vector<AreaTap> xt, yt;
    std::vector<int> xo, yo;
    build_area_tab(w, dw, sx, xt, xo);
    build_area_tab(h, dh, sy, yt, yo);
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c)
                dst[((size_t)y * dw + x) * cn + c] = resize_area_sample(src, (size_t)w * cn, cn, c, y, x, isx, isy, fast, xt.data(), xo.data(), yt.data(), yo.data());
}
void hc_roi(int cols, int rows, int enabled, float rx, float ry, float rw, float rh, int* out4) {
    preprocess_roi(cols, rows, enabled != 0, rx, ry, rw, rh, out4[0], out4[1], out4[2], out4[3]);
}
}
