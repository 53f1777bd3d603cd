// Motion (Laplace) kernels for sm_100a: one fused kernel per pyramid level and direction.
//
//   ingest_down : u8 BGR/gray -> f32 (Lab via OpenCV's LUT) -> pyrDown            (MagnifyCore.hpp:87-93, SpatialFilter.cpp:31)
//   level       : pyrDown + pyrUp + subtract + dual-EMA state update + gain        (SpatialFilter.cpp:25-38, TemporalFilter.cpp:9-22, MagnifyCore.hpp:127-134)
//   collapse    : pyrUp + add                                                      (SpatialFilter.cpp:52-61)
//   egress      : pyrUp + chroma attenuation + input+motion + Lab2BGR + u8         (MagnifyCore.hpp:136-158)
//
// All are HBM-bandwidth-bound stencil / pointwise kernels (no tensor cores).  Tiles are staged in
// shared memory; the row pass of each separable 5-tap filter runs out of that tile.
#include "mc_internal.h"

namespace mc {

namespace {

constexpr float kInv256 = 1.0f / 256.0f;
constexpr float kInv64 = 1.0f / 64.0f;

__device__ __forceinline__ float down5(float a, float b, float c, float d, float e) {
    // cv::pyrDown row/column pass: c*6 + (b+d)*4 + a + e
    return c * 6.0f + (b + d) * 4.0f + a + e;
}

// ------------------------------------------------------------------------------------------------
// ingest_down: tile = 32x16 G1 pixels per CTA, all C channels (the Lab conversion is shared).
// ------------------------------------------------------------------------------------------------
constexpr int ID_TW = 32, ID_TH = 16;
constexpr int ID_SW = 2 * ID_TW + 3, ID_SH = 2 * ID_TH + 3;  // 67 x 35 source pixels
constexpr int ID_SP = ID_SW + 1;                              // smem pitch 68

template <int C>
__global__ void __launch_bounds__(256) k_ingest_down(const uint8_t* __restrict__ in, size_t in_step,
                                                     size_t in_lane_stride, int w0, int h0,
                                                     const LabLutEntry* __restrict__ lut, float* __restrict__ g1,
                                                     int w1, int h1, int pitch1, size_t plane1,
                                                     float* __restrict__ g0, int pitch0, size_t plane0) {
    __shared__ float s0[C][ID_SH][ID_SP];
    __shared__ float sh[C][ID_SH][ID_TW];
    const int lane = blockIdx.z;
    const int x1_0 = blockIdx.x * ID_TW, y1_0 = blockIdx.y * ID_TH;
    const uint8_t* src = in + (size_t)lane * in_lane_stride;
    const int gx0 = 2 * x1_0 - 2, gy0 = 2 * y1_0 - 2;

    for (int idx = threadIdx.x; idx < ID_SH * ID_SW; idx += blockDim.x) {
        const int r = idx / ID_SW, c = idx - r * ID_SW;
        const int gy = reflect101(gy0 + r, h0), gx = reflect101(gx0 + c, w0);
        const uint8_t* p = src + (size_t)gy * in_step + (size_t)gx * C;
        if (C == 3) {
            float L, A, B;
            bgr_u8_to_lab(__ldg(p), __ldg(p + 1), __ldg(p + 2), lut, L, A, B);
            s0[0][r][c] = L;
            s0[C > 1 ? 1 : 0][r][c] = A;
            s0[C > 2 ? 2 : 0][r][c] = B;
            if (g0) {
                // materialise G0 (faithful mode only): each in-image pixel is written by the tile
                // whose interior covers it
                const int ry = gy0 + r, rx = gx0 + c;
                if (ry >= 2 * y1_0 && ry < 2 * (y1_0 + ID_TH) && ry < h0 && rx >= 2 * x1_0 &&
                    rx < 2 * (x1_0 + ID_TW) && rx < w0) {
                    float* q = g0 + (size_t)(lane * C) * plane0 + (size_t)ry * pitch0 + rx;
                    q[0] = L; q[plane0] = A; q[2 * plane0] = B;
                }
            }
        } else {
            const float v = u8_to_unit(__ldg(p));
            s0[0][r][c] = v;
            if (g0) {
                const int ry = gy0 + r, rx = gx0 + c;
                if (ry >= 2 * y1_0 && ry < 2 * (y1_0 + ID_TH) && ry < h0 && rx >= 2 * x1_0 &&
                    rx < 2 * (x1_0 + ID_TW) && rx < w0)
                    g0[(size_t)lane * plane0 + (size_t)ry * pitch0 + rx] = v;
            }
        }
    }
    __syncthreads();
    // row pass
    for (int idx = threadIdx.x; idx < ID_SH * ID_TW; idx += blockDim.x) {
        const int r = idx / ID_TW, x = idx - r * ID_TW;
        const int c = 2 * x + 2;
#pragma unroll
        for (int ch = 0; ch < C; ++ch)
            sh[ch][r][x] = down5(s0[ch][r][c - 2], s0[ch][r][c - 1], s0[ch][r][c], s0[ch][r][c + 1], s0[ch][r][c + 2]);
    }
    __syncthreads();
    // column pass + store
    for (int idx = threadIdx.x; idx < ID_TH * ID_TW; idx += blockDim.x) {
        const int y = idx / ID_TW, x = idx - y * ID_TW;
        const int oy = y1_0 + y, ox = x1_0 + x;
        if (oy < h1 && ox < w1) {
            const int r = 2 * y + 2;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const float v = down5(sh[ch][r - 2][x], sh[ch][r - 1][x], sh[ch][r][x], sh[ch][r + 1][x], sh[ch][r + 2][x]) * kInv256;
                g1[(size_t)(lane * C + ch) * plane1 + (size_t)oy * pitch1 + ox] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// level: tile = 64x16 fine pixels per CTA, one plane per blockIdx.z.
// ------------------------------------------------------------------------------------------------
constexpr int LV_TW = 64, LV_TH = 16;
constexpr int LV_GW = LV_TW + 7, LV_GH = LV_TH + 7;   // fine window 71 x 23 (origin x0-4, y0-4)
constexpr int LV_GP = LV_GW + 1;                      // 72
constexpr int LV_DW = LV_TW / 2 + 2, LV_DH = LV_TH / 2 + 2;  // coarse window 34 x 10 (origin x0/2-1)

__global__ void __launch_bounds__(256) k_level(LevelArgs a) {
    __shared__ float sG[LV_GH][LV_GP];
    __shared__ float sH[LV_GH][LV_DW];
    __shared__ float sD[LV_DH][LV_DW];
    __shared__ float sU[LV_DH][LV_TW];
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * LV_TW, y0 = blockIdx.y * LV_TH;
    const int wf = a.lf.w, hf = a.lf.h, wc = a.lc.w, hc = a.lc.h;
    const float* __restrict__ g = a.g + (size_t)plane * a.lf.plane;

    for (int idx = threadIdx.x; idx < LV_GH * LV_GW; idx += blockDim.x) {
        const int r = idx / LV_GW, c = idx - r * LV_GW;
        const int gy = reflect101(y0 - 4 + r, hf), gx = reflect101(x0 - 4 + c, wf);
        sG[r][c] = __ldg(g + (size_t)gy * a.lf.pitch + gx);
    }
    __syncthreads();
    // pyrDown row pass at the (border-mapped) coarse columns of the window
    for (int idx = threadIdx.x; idx < LV_GH * LV_DW; idx += blockDim.x) {
        const int r = idx / LV_DW, j = idx - r * LV_DW;
        const int im = upsrc(x0 / 2 - 1 + j, wc);
        int c = 2 * im - x0 + 4;
        c = c < 2 ? 2 : (c > LV_GW - 3 ? LV_GW - 3 : c);
        sH[r][j] = down5(sG[r][c - 2], sG[r][c - 1], sG[r][c], sG[r][c + 1], sG[r][c + 2]);
    }
    __syncthreads();
    // pyrDown column pass -> coarse window D (with pyrUp's border rule already applied), store G_{l+1}
    float* __restrict__ gn = a.g_next + (size_t)plane * a.lc.plane;
    for (int idx = threadIdx.x; idx < LV_DH * LV_DW; idx += blockDim.x) {
        const int k = idx / LV_DW, j = idx - k * LV_DW;
        const int iy = y0 / 2 - 1 + k, ix = x0 / 2 - 1 + j;
        const int imy = upsrc(iy, hc);
        int r = 2 * imy - y0 + 4;
        r = r < 2 ? 2 : (r > LV_GH - 3 ? LV_GH - 3 : r);
        const float v = down5(sH[r - 2][j], sH[r - 1][j], sH[r][j], sH[r + 1][j], sH[r + 2][j]) * kInv256;
        sD[k][j] = v;
        if (k >= 1 && k <= LV_TH / 2 && j >= 1 && j <= LV_TW / 2 && iy < hc && ix < wc)
            gn[(size_t)iy * a.lc.pitch + ix] = v;
    }
    __syncthreads();
    // pyrUp row pass
    for (int idx = threadIdx.x; idx < LV_DH * LV_TW; idx += blockDim.x) {
        const int k = idx / LV_TW, x = idx - k * LV_TW;
        const int j0 = (x >> 1) + 1;
        sU[k][x] = (x & 1) ? (sD[k][j0] + sD[k][j0 + 1]) * 4.0f
                           : (sD[k][j0 - 1] + sD[k][j0] * 6.0f + sD[k][j0 + 1]);
    }
    __syncthreads();
    // pyrUp column pass, band, temporal filter
    float* __restrict__ hi = a.hi + (size_t)plane * a.lf.plane;
    float* __restrict__ lo = a.lo + (size_t)plane * a.lf.plane;
    float* __restrict__ m = a.m ? a.m + (size_t)plane * a.lf.plane : nullptr;
    for (int idx = threadIdx.x; idx < LV_TH * LV_TW; idx += blockDim.x) {
        const int y = idx / LV_TW, x = idx - y * LV_TW;
        const int gy = y0 + y, gx = x0 + x;
        if (gy >= hf || gx >= wf) continue;
        const int k0 = (y >> 1) + 1;
        const float up = (y & 1) ? ((sU[k0][x] + sU[k0 + 1][x]) * 4.0f) * kInv64
                                 : (sU[k0 - 1][x] + sU[k0][x] * 6.0f + sU[k0 + 1][x]) * kInv64;
        const float band = sG[y + 4][x + 4] - up;
        const size_t o = (size_t)gy * a.lf.pitch + gx;
        if (a.first) {
            hi[o] = band;
            lo[o] = band;
        } else {
            const float nh = ema(hi[o], band, a.one_minus_c_hi, a.c_hi);
            const float nl = ema(lo[o], band, a.one_minus_c_lo, a.c_lo);
            hi[o] = nh;
            lo[o] = nl;
            if (m) m[o] = (nh - nl) * a.gain;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// collapse: cur_l = pyrUp(cur_{l+1}) + m_l  (in place in m_l)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_coarse_window(float (*sD)[LV_DW], const float* __restrict__ src, int pitch,
                                                   int wc, int hc, int x0, int y0) {
    for (int idx = threadIdx.x; idx < LV_DH * LV_DW; idx += blockDim.x) {
        const int k = idx / LV_DW, j = idx - k * LV_DW;
        const int iy = upsrc(y0 / 2 - 1 + k, hc), ix = upsrc(x0 / 2 - 1 + j, wc);
        sD[k][j] = __ldg(src + (size_t)iy * pitch + ix);
    }
}

__device__ __forceinline__ void up_rows(float (*sU)[LV_TW], const float (*sD)[LV_DW]) {
    for (int idx = threadIdx.x; idx < LV_DH * LV_TW; idx += blockDim.x) {
        const int k = idx / LV_TW, x = idx - k * LV_TW;
        const int j0 = (x >> 1) + 1;
        sU[k][x] = (x & 1) ? (sD[k][j0] + sD[k][j0 + 1]) * 4.0f
                           : (sD[k][j0 - 1] + sD[k][j0] * 6.0f + sD[k][j0 + 1]);
    }
}

__device__ __forceinline__ float up_col(const float (*sU)[LV_TW], int y, int x) {
    const int k0 = (y >> 1) + 1;
    return (y & 1) ? ((sU[k0][x] + sU[k0 + 1][x]) * 4.0f) * kInv64
                   : (sU[k0 - 1][x] + sU[k0][x] * 6.0f + sU[k0 + 1][x]) * kInv64;
}

__global__ void __launch_bounds__(256) k_collapse(Level lf, Level lc, float* __restrict__ m_fine,
                                                  const float* __restrict__ m_coarse) {
    __shared__ float sD[LV_DH][LV_DW];
    __shared__ float sU[LV_DH][LV_TW];
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * LV_TW, y0 = blockIdx.y * LV_TH;
    load_coarse_window(sD, m_coarse + (size_t)plane * lc.plane, lc.pitch, lc.w, lc.h, x0, y0);
    __syncthreads();
    up_rows(sU, sD);
    __syncthreads();
    float* __restrict__ mf = m_fine + (size_t)plane * lf.plane;
    for (int idx = threadIdx.x; idx < LV_TH * LV_TW; idx += blockDim.x) {
        const int y = idx / LV_TW, x = idx - y * LV_TW;
        const int gy = y0 + y, gx = x0 + x;
        if (gy >= lf.h || gx >= lf.w) continue;
        const size_t o = (size_t)gy * lf.pitch + gx;
        mf[o] = up_col(sU, y, x) + mf[o];
    }
}

// ------------------------------------------------------------------------------------------------
// egress: tile = 64x16 output pixels, all channels.
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) k_egress(const uint8_t* __restrict__ in, size_t in_step, size_t in_lane_stride,
                                                uint8_t* __restrict__ out, size_t out_step, size_t out_lane_stride,
                                                int w0, int h0, const LabLutEntry* __restrict__ lut,
                                                const float4* __restrict__ gtab, LabInvCoeffs coeffs,
                                                const float* __restrict__ m1, Level l1, float chroma,
                                                float* __restrict__ fout) {
    __shared__ float sD[C][LV_DH][LV_DW];
    __shared__ float sU[C][LV_DH][LV_TW];
    const int lane = blockIdx.z;
    const int x0 = blockIdx.x * LV_TW, y0 = blockIdx.y * LV_TH;
    if (m1) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch)
            load_coarse_window(sD[ch], m1 + (size_t)(lane * C + ch) * l1.plane, l1.pitch, l1.w, l1.h, x0, y0);
        __syncthreads();
#pragma unroll
        for (int ch = 0; ch < C; ++ch) up_rows(sU[ch], sD[ch]);
        __syncthreads();
    }
    const uint8_t* src = in + (size_t)lane * in_lane_stride;
    uint8_t* dst = out + (size_t)lane * out_lane_stride;
    for (int idx = threadIdx.x; idx < LV_TH * LV_TW; idx += blockDim.x) {
        const int y = idx / LV_TW, x = idx - y * LV_TW;
        const int gy = y0 + y, gx = x0 + x;
        if (gy >= h0 || gx >= w0) continue;
        const uint8_t* p = src + (size_t)gy * in_step + (size_t)gx * C;
        uint8_t* q = dst + (size_t)gy * out_step + (size_t)gx * C;
        if (C == 3) {
            float L, A, B;
            bgr_u8_to_lab(__ldg(p), __ldg(p + 1), __ldg(p + 2), lut, L, A, B);
            if (m1) {
                // motion planes a,b *= chromAttenuation, then output = input + motion (MagnifyCore.hpp:140-148)
                L = L + up_col(sU[0], y, x);
                A = A + up_col(sU[C > 1 ? 1 : 0], y, x) * chroma;
                B = B + up_col(sU[C > 2 ? 2 : 0], y, x) * chroma;
            }
            float ob, og, orr;
            lab_to_bgr(L, A, B, coeffs, gtab, ob, og, orr);
            q[0] = unit_to_u8(ob);
            q[1] = unit_to_u8(og);
            q[2] = unit_to_u8(orr);
            if (fout) {
                float* f = fout + (((size_t)lane * h0 + gy) * w0 + gx) * 3;
                f[0] = ob; f[1] = og; f[2] = orr;
            }
        } else {
            float v = u8_to_unit(__ldg(p));
            if (m1) v = v + up_col(sU[0], y, x);
            q[0] = unit_to_u8(v);
            if (fout) fout[((size_t)lane * h0 + gy) * w0 + gx] = v;
        }
    }
}

__global__ void k_copy(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

inline unsigned cdiv(int a, int b) { return (unsigned)((a + b - 1) / b); }

}  // namespace

cudaError_t launch_ingest_down(const FrameIO& io, const DeviceTables& tb, const Level& l0, const Level& l1,
                               float* g1, float* g0, cudaStream_t s) {
    dim3 grid(cdiv(l1.w, ID_TW), cdiv(l1.h, ID_TH), io.lanes);
    if (io.channels == 3)
        k_ingest_down<3><<<grid, 256, 0, s>>>(io.in, io.in_step, io.in_lane_stride, l0.w, l0.h, tb.lab_lut, g1, l1.w,
                                              l1.h, l1.pitch, l1.plane, g0, l0.pitch, l0.plane);
    else
        k_ingest_down<1><<<grid, 256, 0, s>>>(io.in, io.in_step, io.in_lane_stride, l0.w, l0.h, tb.lab_lut, g1, l1.w,
                                              l1.h, l1.pitch, l1.plane, g0, l0.pitch, l0.plane);
    return cudaGetLastError();
}

cudaError_t launch_level(const LevelArgs& a, cudaStream_t s) {
    dim3 grid(cdiv(a.lf.w, LV_TW), cdiv(a.lf.h, LV_TH), a.planes);
    k_level<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_collapse(const Level& lf, const Level& lc, float* m_fine, const float* m_coarse, int planes,
                            cudaStream_t s) {
    dim3 grid(cdiv(lf.w, LV_TW), cdiv(lf.h, LV_TH), planes);
    k_collapse<<<grid, 256, 0, s>>>(lf, lc, m_fine, m_coarse);
    return cudaGetLastError();
}

cudaError_t launch_egress(const FrameIO& io, const DeviceTables& tb, const Level& l0, const Level& l1,
                          const float* m1, float chroma, float* fout, cudaStream_t s) {
    dim3 grid(cdiv(l0.w, LV_TW), cdiv(l0.h, LV_TH), io.lanes);
    if (io.channels == 3)
        k_egress<3><<<grid, 256, 0, s>>>(io.in, io.in_step, io.in_lane_stride, io.out, io.out_step, io.out_lane_stride,
                                         l0.w, l0.h, tb.lab_lut, tb.inv_gamma, tb.inv_coeffs, m1, l1, chroma, fout);
    else
        k_egress<1><<<grid, 256, 0, s>>>(io.in, io.in_step, io.in_lane_stride, io.out, io.out_step, io.out_lane_stride,
                                         l0.w, l0.h, tb.lab_lut, tb.inv_gamma, tb.inv_coeffs, m1, l1, chroma, fout);
    return cudaGetLastError();
}

cudaError_t launch_copy_planes(float* dst, const float* src, size_t n, cudaStream_t s) {
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks == 0) blocks = 1;
    k_copy<<<blocks, 256, 0, s>>>(dst, src, n);
    return cudaGetLastError();
}

}  // namespace mc
