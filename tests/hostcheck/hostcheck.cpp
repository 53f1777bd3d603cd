// TEST INFRASTRUCTURE ONLY.  Compiles the product's __host__ __device__ per-pixel functions
// (csrc/mc_math.cuh) and host tables (csrc/mc_tables.cpp) for the CPU so the non-GPU test-suite can
// compare them with the oracle.  Never loaded by the product; the product has no CPU path.
#include <cstdint>
#include <vector>

#include "mc_internal.h"

using namespace mc;

namespace {
struct Tables {
    std::vector<LabLutEntry> lut;
    std::vector<float4> gam;
    LabInvCoeffs k;
    Tables() {
        build_lab_lut_packed(lut);
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
}
