// Per-pixel arithmetic shared by the kernels.  Everything here is __host__ __device__ so the CPU
// test-suite can compile the *same* functions into a check library (tests/hostcheck) and compare
// them with the oracle without a GPU; the product never calls them on the host.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif

namespace mc {

constexpr int kLabLutDim = 33;          // OpenCV LAB_LUT_DIM
constexpr int kGammaTabSize = 1024;     // OpenCV GAMMA_TAB_SIZE

// One packed LUT entry: Lab int16 triples of lattice points (b,g,r) and (b,g,r+1).
struct alignas(16) LabLutEntry { int16_t v[8]; };  // {L0,a0,b0,L1,a1,b1,0,0}

MC_HD int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// pyrUp source-index rule (cv::pyrUp): s[-1] := s[1], s[n] := s[n-1].
MC_HD int upsrc(int i, int n) {
    if (i < 0) return 1;
    if (i >= n) return n - 1;
    return i;
}

// u8 -> f32 as Mat::convertTo(CV_32F, 1/255) does it (MagnifyCore.hpp:89): x * (float)(1/255).
MC_HD float u8_to_unit(uint8_t v) { return (float)v * 0.003921568859368563f; }

// f32 -> u8 as Mat::convertTo(CV_8U, 255, 1/255) does it (MagnifyCore.hpp:153):
// saturate(round_half_even(fma(x, 255, (float)(1/255)))).
MC_HD uint8_t unit_to_u8(float x) {
    float v = fmaf(x, 255.0f, 0.003921568859368563f);
    v = rintf(v);
    if (!(v > 0.0f)) return 0;   // also NaN -> 0
    if (v > 255.0f) return 255;
    return (uint8_t)(int)v;
}

// generic convertTo(CV_8U, alpha, beta) used by Color egress (MagnifyCore.hpp:202-203)
MC_HD uint8_t scaled_to_u8(float x, float a, float b) {
    float v = rintf(fmaf(x, a, b));
    if (!(v > 0.0f)) return 0;
    if (v > 255.0f) return 255;
    return (uint8_t)(int)v;
}

#if defined(__CUDACC__)
#define MC_LDG16(p) __ldg(reinterpret_cast<const int4*>(p))
#endif

// cv::cvtColor(COLOR_BGR2Lab) on CV_32F input that came from u8/255 — bit-exact restatement of
// OpenCV's 33^3 int16 LUT + 4-bit fixed-point trilinear interpolation.
// lut is [b][g][r] packed entries.  Fixed-point result: L*2^14/100, (a+128)*64, (b+128)*64.
MC_HD void bgr_u8_to_lab_fixed(uint8_t b8, uint8_t g8, uint8_t r8, const LabLutEntry* __restrict__ lut,
                               int& sL, int& sA, int& sB) {
#if defined(__CUDA_ARCH__)
    const int cb = __float2int_rn(u8_to_unit(b8) * 16384.0f);
    const int cg = __float2int_rn(u8_to_unit(g8) * 16384.0f);
    const int cr = __float2int_rn(u8_to_unit(r8) * 16384.0f);
#else
    const int cb = (int)lrintf(u8_to_unit(b8) * 16384.0f);
    const int cg = (int)lrintf(u8_to_unit(g8) * 16384.0f);
    const int cr = (int)lrintf(u8_to_unit(r8) * 16384.0f);
#endif
    const int tb = cb >> 9, tg = cg >> 9, tr = cr >> 9;
    const int xb = (cb >> 5) & 15, xg = (cg >> 5) & 15, xr = (cr >> 5) & 15;
    const int tb1 = tb + 1 > 32 ? 32 : tb + 1, tg1 = tg + 1 > 32 ? 32 : tg + 1;
    sL = 0; sA = 0; sB = 0;
#pragma unroll
    for (int db = 0; db < 2; ++db) {
#pragma unroll
        for (int dg = 0; dg < 2; ++dg) {
            const int ib = db ? tb1 : tb, ig = dg ? tg1 : tg;
            const int wbg = (db ? xb : 16 - xb) * (dg ? xg : 16 - xg);
            const LabLutEntry* e = lut + ((ib * kLabLutDim + ig) * kLabLutDim + tr);
#if defined(__CUDA_ARCH__)
            const int4 q = MC_LDG16(e);
            const int L0 = (int)(short)(q.x & 0xffff), a0 = q.x >> 16;
            const int b0 = (int)(short)(q.y & 0xffff), L1 = q.y >> 16;
            const int a1 = (int)(short)(q.z & 0xffff), b1 = q.z >> 16;
#else
            const int L0 = e->v[0], a0 = e->v[1], b0 = e->v[2], L1 = e->v[3], a1 = e->v[4], b1 = e->v[5];
#endif
            sL += wbg * ((16 - xr) * L0 + xr * L1);
            sA += wbg * ((16 - xr) * a0 + xr * a1);
            sB += wbg * ((16 - xr) * b0 + xr * b1);
        }
    }
    sL = (sL + 2048) >> 12;
    sA = (sA + 2048) >> 12;
    sB = (sB + 2048) >> 12;
}

// Float Lab as OpenCV returns it: L in [0,100], a,b in [-128,128).
MC_HD void bgr_u8_to_lab(uint8_t b8, uint8_t g8, uint8_t r8, const LabLutEntry* __restrict__ lut,
                         float& L, float& A, float& B) {
    int sL, sA, sB;
    bgr_u8_to_lab_fixed(b8, g8, r8, lut, sL, sA, sB);
    L = (float)sL * (100.0f / 16384.0f);
    A = fmaf((float)sA, 1.0f / 64.0f, -128.0f);
    B = fmaf((float)sB, 1.0f / 64.0f, -128.0f);
}

// Coefficients for Lab -> BGR (XYZ2sRGB_D65 rows scaled by the D65 white point), filled on the host.
struct LabInvCoeffs { float c[9]; };  // rows: B, G, R ; columns: X, Y, Z

MC_HD float spline_gamma(float v, const float4* __restrict__ tab) {
    // OpenCV splineInterpolate(x*GAMMA_TAB_SIZE, sRGBInvGammaTab, GAMMA_TAB_SIZE)
    const float xs = v * (float)kGammaTabSize;
    int ix = (int)xs;
    ix = ix < 0 ? 0 : (ix > kGammaTabSize - 1 ? kGammaTabSize - 1 : ix);
    const float fr = xs - (float)ix;
#if defined(__CUDA_ARCH__)
    const float4 t = __ldg(tab + ix);
#else
    const float4 t = tab[ix];
#endif
    return ((t.w * fr + t.z) * fr + t.y) * fr + t.x;
}

// cv::cvtColor(COLOR_Lab2BGR) on CV_32F (analytic inverse + spline-interpolated sRGB gamma,
// output clipped to [0,1]); restated from OpenCV's Lab2RGBfloat, checked against cv2 to ~1e-5.
MC_HD void lab_to_bgr(float L, float a, float b, const LabInvCoeffs& k, const float4* __restrict__ gtab,
                      float& ob, float& og, float& orr) {
    // constant divisions are written as reciprocal multiplies (<= 1 ulp from the divide; the parity
    // budget is 1e-4) — IEEE divides would triple the instruction count of the egress kernel
    float Y, fy;
    if (L <= 8.0f) {
        Y = L * (1.0f / 903.3f);
        fy = 7.787f * Y + 16.0f / 116.0f;
    } else {
        fy = (L + 16.0f) * (1.0f / 116.0f);
        Y = fy * fy * fy;
    }
    float fx = a * (1.0f / 500.0f) + fy;
    float fz = fy - b * (1.0f / 200.0f);
    const float fth = 6.0f / 29.0f;
    const float X = fx <= fth ? (fx - 16.0f / 116.0f) * (1.0f / 7.787f) : fx * fx * fx;
    const float Z = fz <= fth ? (fz - 16.0f / 116.0f) * (1.0f / 7.787f) : fz * fz * fz;
    float vb = k.c[0] * X + k.c[1] * Y + k.c[2] * Z;
    float vg = k.c[3] * X + k.c[4] * Y + k.c[5] * Z;
    float vr = k.c[6] * X + k.c[7] * Y + k.c[8] * Z;
    vb = fminf(fmaxf(vb, 0.0f), 1.0f);
    vg = fminf(fmaxf(vg, 0.0f), 1.0f);
    vr = fminf(fmaxf(vr, 0.0f), 1.0f);
    ob = spline_gamma(vb, gtab);
    og = spline_gamma(vg, gtab);
    orr = spline_gamma(vr, gtab);
}

// iirFilter (TemporalFilter.cpp:9-22): cv::addWeighted rounds once from a double sum.
MC_HD float ema(float state, float x, double one_minus_c, double c) {
    return (float)((double)state * one_minus_c + (double)x * c);
}

}  // namespace mc
