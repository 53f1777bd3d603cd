// Phase (Riesz) mode — device twin of magcore::magnifyRiesz (reference
// src/processing/magnification/MagnifyCore.hpp:209-279) with RieszPyramid.cpp / TemporalFilter.cpp:299-362.
//
//   analysis : 9x9 high-pass band + 9x9 2*low-pass, subsampled            (RieszPyramid.cpp:215-238)
//   phase    : Riesz pair (1x5 / 5x1) + quaternionic phase difference +
//              amplitude + both 2nd-order Butterworth IIRs (DF-II)         (RieszPyramid.cpp:66-111, TemporalFilter.cpp:340-351)
//   amplify  : separable 13-tap Gaussian of (A, cos, sin) + phase shift    (RieszPyramid.cpp:114-144)
//   collapse : zero-injected 2*LP up-sampling + HP of the band, summed     (RieszPyramid.cpp:304-325)
//   egress   : merge L' with a,b -> Lab2BGR -> u8                           (MagnifyCore.hpp:272-276)
//
// The point-wise quaternion algebra uses explicit round-to-nearest intrinsics (no FMA contraction)
// so that, given identical inputs, it rounds exactly like OpenCV's element-wise cv::multiply / add /
// divide / sqrt calls — acos near 1 is ill-conditioned.
#include <cmath>
#include <cstring>

#include "mc_modes.h"
#include "mc_tma.cuh"

namespace mc {

namespace {

// literal 4-decimal tap tables of the Riesz-pyramid paper, as the reference writes them
// (RieszPyramid.cpp:146-167); the low-pass is used x2 on analysis and synthesis (:232, :316).
__constant__ float c_hp[81] = {
    0.0000f, 0.0003f, 0.0011f, 0.0022f, 0.0027f, 0.0022f, 0.0011f, 0.0003f, 0.0000f,
    0.0003f, 0.0020f, 0.0059f, 0.0103f, 0.0123f, 0.0103f, 0.0059f, 0.0020f, 0.0003f,
    0.0011f, 0.0059f, 0.0151f, 0.0249f, 0.0292f, 0.0249f, 0.0151f, 0.0059f, 0.0011f,
    0.0022f, 0.0103f, 0.0249f, 0.0402f, 0.0469f, 0.0402f, 0.0249f, 0.0103f, 0.0022f,
    0.0027f, 0.0123f, 0.0292f, 0.0469f, -0.9455f, 0.0469f, 0.0292f, 0.0123f, 0.0027f,
    0.0022f, 0.0103f, 0.0249f, 0.0402f, 0.0469f, 0.0402f, 0.0249f, 0.0103f, 0.0022f,
    0.0011f, 0.0059f, 0.0151f, 0.0249f, 0.0292f, 0.0249f, 0.0151f, 0.0059f, 0.0011f,
    0.0003f, 0.0020f, 0.0059f, 0.0103f, 0.0123f, 0.0103f, 0.0059f, 0.0020f, 0.0003f,
    0.0000f, 0.0003f, 0.0011f, 0.0022f, 0.0027f, 0.0022f, 0.0011f, 0.0003f, 0.0000f,
};
__constant__ float c_lp[81] = {
    -0.0001f, -0.0007f, -0.0023f, -0.0046f, -0.0057f, -0.0046f, -0.0023f, -0.0007f, -0.0001f,
    -0.0007f, -0.0030f, -0.0047f, -0.0025f, -0.0003f, -0.0025f, -0.0047f, -0.0030f, -0.0007f,
    -0.0023f, -0.0047f, 0.0054f, 0.0272f, 0.0387f, 0.0272f, 0.0054f, -0.0047f, -0.0023f,
    -0.0046f, -0.0025f, 0.0272f, 0.0706f, 0.0910f, 0.0706f, 0.0272f, -0.0025f, -0.0046f,
    -0.0057f, -0.0003f, 0.0387f, 0.0910f, 0.1138f, 0.0910f, 0.0387f, -0.0003f, -0.0057f,
    -0.0046f, -0.0025f, 0.0272f, 0.0706f, 0.0910f, 0.0706f, 0.0272f, -0.0025f, -0.0046f,
    -0.0023f, -0.0047f, 0.0054f, 0.0272f, 0.0387f, 0.0272f, 0.0054f, -0.0047f, -0.0023f,
    -0.0007f, -0.0030f, -0.0047f, -0.0025f, -0.0003f, -0.0025f, -0.0047f, -0.0030f, -0.0007f,
    -0.0001f, -0.0007f, -0.0023f, -0.0046f, -0.0057f, -0.0046f, -0.0023f, -0.0007f, -0.0001f,
};

constexpr int RT_W = 32, RT_H = 16;                 // output tile

__device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subr(float a, float b) { return __fsub_rn(a, b); }

// cv::filter2D's f32 engine (FilterVec_32f + scalar remainder; OpenCV 4.x, the AVX2 dispatch both this container and the
// B200 boxes run, tools/probe_filter2d_order.py): every output accumulates its non-zero taps in raster order — with one
// FMA per tap in the vectorised columns x < (w & ~7), with multiply-then-add in the scalar tail columns.  The device
// kernels follow that rule column by column, which makes the band planes and the Riesz pair bit-identical to the
// reference for every width (acos near 1 turns a last-ulp difference here into visible differences downstream).
__device__ __forceinline__ float f2d(float c, float v, float acc, bool tail) {
    return tail ? __fadd_rn(acc, __fmul_rn(c, v)) : fmaf(c, v, acc);
}

// 9x9 kernels are register-blocked: a thread owns a 1x4 strip, loads 12 tile values per kernel row with three
// 128-bit shared-memory reads and issues 36 FMAs on them.  Rows are dealt to warps so that both rows of a
// warp have the same parity (the sub-sampled low-pass / the zero-injected up-sampling only touch one parity),
// which keeps every parity branch warp-uniform.
constexpr int R9_W = 64, R9_H = 16;                  // output tile
constexpr int R9_SW = R9_W + 8, R9_SH = R9_H + 8;    // + 4 halo

__device__ __forceinline__ int r9_row(int tid) {     // tile row of this thread: warp w -> rows {b, b+2}, b = 4(w>>1) + (w&1)
    const int w = tid >> 5, half = (tid >> 4) & 1;
    return 4 * (w >> 1) + (w & 1) + 2 * half;
}

// Stages the (R9_SH x R9_SW) window of one plane, origin (x0-4, y0-4), BORDER_REFLECT_101.  Interior tiles with a
// tensor map: ONE cp.async.bulk.tensor copy issued by one thread, completion on an mbarrier (`bar` must have been
// initialised and fenced by the caller before the block-wide barrier that precedes this call); otherwise 128-bit loads
// (interior) or scalar reflected loads (tiles touching an image border).
__device__ __forceinline__ void r9_load_tile(float (*s)[R9_SW], const float* __restrict__ src, const Level& l, int x0, int y0,
                                             const CUtensorMap* tm = nullptr, int plane = 0, uint64_t* bar = nullptr) {
    const bool interior = x0 >= 4 && x0 + R9_W + 4 <= l.w && y0 >= 4 && y0 + R9_H + 4 <= l.h;
    if (interior && tm) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, R9_SH * R9_SW * sizeof(float));
            tma_load_3d(&s[0][0], tm, x0 - 4, y0 - 4, plane, bar);
        }
        mbar_wait(bar, 0);
    } else if (interior) {
        for (int i = threadIdx.x; i < R9_SH * (R9_SW / 4); i += 256) {
            const int r = i / (R9_SW / 4), c4 = i - r * (R9_SW / 4);
            *reinterpret_cast<float4*>(&s[r][4 * c4]) =
                __ldg(reinterpret_cast<const float4*>(src + (size_t)(y0 - 4 + r) * l.pitch + (x0 - 4 + 4 * c4)));
        }
    } else {
        for (int i = threadIdx.x; i < R9_SH * R9_SW; i += 256) {
            const int r = i / R9_SW, c = i - r * R9_SW;
            s[r][c] = __ldg(src + (size_t)reflect101(y0 - 4 + r, l.h) * l.pitch + reflect101(x0 - 4 + c, l.w));
        }
    }
}

// hp = filter2D(oct, HP) ; next = subsample(filter2D(oct, 2*LP))   (REFLECT_101, correlation)
__global__ void __launch_bounds__(256) k_riesz_analysis(Level l, Level ln, const float* __restrict__ oct,
                                                        float* __restrict__ hp, float* __restrict__ next,
                                                        const __grid_constant__ CUtensorMap tm, int use_tma) {
    __shared__ __align__(128) float s[R9_SH][R9_SW];
    __shared__ __align__(8) uint64_t bar;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * R9_W, y0 = blockIdx.y * R9_H;
    if (use_tma) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
    }
    r9_load_tile(s, oct + (size_t)plane * l.plane, l, x0, y0, use_tma ? &tm : nullptr, plane, &bar);
    __syncthreads();
    const int tx = threadIdx.x & 15, y = r9_row(threadIdx.x);
    const int gy = y0 + y, gx = x0 + 4 * tx;
    if (gy >= l.h || gx >= l.w) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, lp0 = 0.f, lp1 = 0.f;
    const bool even_row = next != nullptr && !(gy & 1);   // warp-uniform
    const int tail_from = (l.w & ~7) - gx;                 // strip columns p >= tail_from are filter2D's scalar-tail columns
    if (tail_from >= 4) {
#pragma unroll
        for (int ky = 0; ky < 9; ++ky) {
            const float4 a0 = *reinterpret_cast<const float4*>(&s[y + ky][4 * tx]);
            const float4 a1 = *reinterpret_cast<const float4*>(&s[y + ky][4 * tx + 4]);
            const float4 a2 = *reinterpret_cast<const float4*>(&s[y + ky][4 * tx + 8]);
            const float v[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
            for (int kx = 0; kx < 9; ++kx) {
                const float c = c_hp[ky * 9 + kx];
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[p] = fmaf(c, v[p + kx], acc[p]);
            }
            if (even_row) {
#pragma unroll
                for (int kx = 0; kx < 9; ++kx) {
                    const float c = 2.0f * c_lp[ky * 9 + kx];
                    lp0 = fmaf(c, v[kx], lp0);
                    lp1 = fmaf(c, v[2 + kx], lp1);
                }
            }
        }
    } else {   // the last strip(s) of a row whose width is not a multiple of 8
#pragma unroll
        for (int ky = 0; ky < 9; ++ky) {
            const float4 a0 = *reinterpret_cast<const float4*>(&s[y + ky][4 * tx]);
            const float4 a1 = *reinterpret_cast<const float4*>(&s[y + ky][4 * tx + 4]);
            const float4 a2 = *reinterpret_cast<const float4*>(&s[y + ky][4 * tx + 8]);
            const float v[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
            for (int kx = 0; kx < 9; ++kx) {
                const float c = c_hp[ky * 9 + kx];
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[p] = f2d(c, v[p + kx], acc[p], p >= tail_from);
            }
            if (even_row) {
#pragma unroll
                for (int kx = 0; kx < 9; ++kx) {
                    const float c = 2.0f * c_lp[ky * 9 + kx];
                    lp0 = f2d(c, v[kx], lp0, 0 >= tail_from);
                    lp1 = f2d(c, v[2 + kx], lp1, 2 >= tail_from);
                }
            }
        }
    }
    float* o = hp + (size_t)plane * l.plane + (size_t)gy * l.pitch + gx;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);   // rows are padded to 32 floats
    if (even_row) {
        float* q = next + (size_t)plane * ln.plane + (size_t)(gy >> 1) * ln.pitch + (gx >> 1);
        *reinterpret_cast<float2*>(q) = make_float2(lp0, lp1);                        // gx/2 is even -> 8-byte aligned; padded row
    }
}

struct Butter { double b0, b1, b2, a1, a2; };

struct PhaseArgs {
    Level l;
    const float* low;                      // this frame's band (hp)
    const float *plow, *prx, *pry;         // prior pyramid (old); null -> use this frame's (cutoff change)
    float *rx, *ry;                        // this frame's Riesz pair (written)
    float *ph_c, *ph_s;                    // accumulated phase
    float *lo_r0c, *lo_r0s, *lo_r1c, *lo_r1s, *hi_r0c, *hi_r0s, *hi_r1c, *hi_r1s;
    float *amp, *t_c, *t_s;                // amplitude, (hiIIR - loIIR) * amplitude
    Butter lo, hi;
};

__device__ __forceinline__ float muld(float x, double s) { return (float)((double)x * s); }  // cv::multiply(Mat, double)

__device__ __forceinline__ float iir_step(float phase, float& r0, float& r1, const Butter& k) {
    // RieszTemporalFilter::IIRTemporalFilter (TemporalFilter.cpp:340-351), Direct Form II
    const float y = addr(muld(phase, k.b0), r0);
    r0 = subr(addr(muld(phase, k.b1), r1), muld(y, k.a1));
    r1 = subr(muld(phase, k.b2), muld(y, k.a2));
    return y;
}

__global__ void __launch_bounds__(256) k_riesz_phase(const PhaseArgs a) {
    __shared__ float s[RT_H + 4][RT_W + 4 + 1];
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * RT_W, y0 = blockIdx.y * RT_H;
    const Level l = a.l;
    const size_t pb = (size_t)plane * l.plane;
    for (int i = threadIdx.x; i < (RT_H + 4) * (RT_W + 4); i += 256) {
        const int r = i / (RT_W + 4), c = i - r * (RT_W + 4);
        s[r][c] = __ldg(a.low + pb + (size_t)reflect101(y0 - 2 + r, l.h) * l.pitch + reflect101(x0 - 2 + c, l.w));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RT_H * RT_W; i += 256) {
        const int y = i / RT_W, x = i - y * RT_W;
        const int gy = y0 + y, gx = x0 + x;
        if (gy >= l.h || gx >= l.w) continue;
        const size_t o = pb + (size_t)gy * l.pitch + gx;
        // RieszPyramidLevel::build (RieszPyramid.cpp:66-78): taps [-0.2, -0.48, 0, 0.48, 0.2]
        const float low = s[y + 2][x + 2];
        const bool tail = gx >= (l.w & ~7);     // filter2D's scalar-tail columns, see f2d()
        float rx = mulr(-0.2f, s[y + 2][x]);
        rx = f2d(-0.48f, s[y + 2][x + 1], rx, tail);
        rx = f2d(0.48f, s[y + 2][x + 3], rx, tail);
        rx = f2d(0.2f, s[y + 2][x + 4], rx, tail);
        float ry = mulr(-0.2f, s[y][x + 2]);
        ry = f2d(-0.48f, s[y + 1][x + 2], ry, tail);
        ry = f2d(0.48f, s[y + 3][x + 2], ry, tail);
        ry = f2d(0.2f, s[y + 4][x + 2], ry, tail);
        a.rx[o] = rx;
        a.ry[o] = ry;
        const float plow = a.plow ? a.plow[o] : low;
        const float prx = a.plow ? a.prx[o] : rx;
        const float pry = a.plow ? a.pry[o] : ry;
        // computePhaseDifferenceAndAmplitude (RieszPyramid.cpp:81-111)
        const float q_real = addr(addr(mulr(low, plow), mulr(rx, prx)), mulr(ry, pry));
        const float neg_low = -low;
        const float qx = addr(mulr(prx, neg_low), mulr(rx, plow));
        const float qy = addr(mulr(pry, neg_low), mulr(ry, plow));
        const float xy_sq = addr(mulr(qx, qx), mulr(qy, qy));
        const float q_amp = __fsqrt_rn(addr(mulr(q_real, q_real), xy_sq));
        const float tmp = __fdiv_rn(q_real, q_amp);
        float phi;
        if (tmp < -1.0f) phi = -1.0f;           // reference quirk: clamps to +-1.0 *radians* (RieszPyramid.cpp:15-18)
        else if (tmp > 1.0f) phi = 1.0f;
        else phi = acosf(tmp);
        const float xy_sqrt = __fsqrt_rn(xy_sq);
        float pd_c = mulr(__fdiv_rn(qx, xy_sqrt), phi);
        float pd_s = mulr(__fdiv_rn(qy, xy_sqrt), phi);
        if (pd_c != pd_c) pd_c = 0.f;           // cv::patchNaNs
        if (pd_s != pd_s) pd_s = 0.f;
        const float amp = __fsqrt_rn(q_amp);
        // temporal band-pass of the accumulated phase (MagnifyCore.hpp:259-264)
        const float ph_c = addr(a.ph_c[o], pd_c), ph_s = addr(a.ph_s[o], pd_s);
        a.ph_c[o] = ph_c;
        a.ph_s[o] = ph_s;
        float r0, r1;
        r0 = a.lo_r0c[o]; r1 = a.lo_r1c[o];
        const float lo_c = iir_step(ph_c, r0, r1, a.lo);
        a.lo_r0c[o] = r0; a.lo_r1c[o] = r1;
        r0 = a.lo_r0s[o]; r1 = a.lo_r1s[o];
        const float lo_s = iir_step(ph_s, r0, r1, a.lo);
        a.lo_r0s[o] = r0; a.lo_r1s[o] = r1;
        r0 = a.hi_r0c[o]; r1 = a.hi_r1c[o];
        const float hi_c = iir_step(ph_c, r0, r1, a.hi);
        a.hi_r0c[o] = r0; a.hi_r1c[o] = r1;
        r0 = a.hi_r0s[o]; r1 = a.hi_r1s[o];
        const float hi_s = iir_step(ph_s, r0, r1, a.hi);
        a.hi_r0s[o] = r0; a.hi_r1s[o] = r1;
        // normalize(): change = highpassIIR - lowpassIIR ; result = change .* amplitude (RieszPyramid.cpp:118-120)
        a.amp[o] = amp;
        a.t_c[o] = mulr(subr(hi_c, lo_c), amp);
        a.t_s[o] = mulr(subr(hi_s, lo_s), amp);
    }
}

struct Gauss13 { float k[13]; };

struct AmpArgs {
    Level l;
    const float *amp, *t_c, *t_s;          // inputs to blur
    const float *low, *rx, *ry;            // this frame's band and Riesz pair
    float* out;                            // amplified band
    Gauss13 g;
    float alpha, thresh;
};

// Tile 64 x 16, thread strip 1 x 4 (same row dealing as the 9x9 kernels is not needed here).  Row pass:
// 16 tile values -> 4 outputs per plane with three/four 128-bit shared loads; column pass likewise from the
// row-pass buffer; then the point-wise amplification on float4 global accesses.
constexpr int GA_TW = 64, GA_TH = 16;
constexpr int GA_W = GA_TW + 16, GA_H = GA_TH + 12;   // window origin (x0-8, y0-6): 8 columns of left halo keep float4 alignment

__device__ __forceinline__ float g13(const Gauss13& g, const float* v) {
    // symmetric 13-tap: centre + pairs (cv::sepFilter2D symmetric row/column filters)
    float acc = g.k[6] * v[6];
#pragma unroll
    for (int j = 1; j <= 6; ++j) acc = fmaf(g.k[6 + j], v[6 - j] + v[6 + j], acc);
    return acc;
}

__global__ void __launch_bounds__(256) k_riesz_amplify(const AmpArgs a) {
    __shared__ __align__(16) float s[3][GA_H][GA_W];
    __shared__ __align__(16) float sr[3][GA_H][GA_TW];
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * GA_TW, y0 = blockIdx.y * GA_TH;
    const Level l = a.l;
    const size_t pb = (size_t)plane * l.plane;
    const float* srcs[3] = {a.amp, a.t_c, a.t_s};
    const bool interior = x0 >= 8 && x0 + GA_TW + 8 <= l.w && y0 >= 6 && y0 + GA_TH + 6 <= l.h;
    if (interior) {
        for (int i = threadIdx.x; i < GA_H * (GA_W / 4); i += 256) {
            const int r = i / (GA_W / 4), c4 = i - r * (GA_W / 4);
            const size_t o = pb + (size_t)(y0 - 6 + r) * l.pitch + (x0 - 8 + 4 * c4);
#pragma unroll
            for (int q = 0; q < 3; ++q) *reinterpret_cast<float4*>(&s[q][r][4 * c4]) = __ldg(reinterpret_cast<const float4*>(srcs[q] + o));
        }
    } else {
        for (int i = threadIdx.x; i < GA_H * GA_W; i += 256) {
            const int r = i / GA_W, c = i - r * GA_W;
            const size_t o = pb + (size_t)reflect101(y0 - 6 + r, l.h) * l.pitch + reflect101(x0 - 8 + c, l.w);
#pragma unroll
            for (int q = 0; q < 3; ++q) s[q][r][c] = __ldg(srcs[q] + o);
        }
    }
    __syncthreads();
    // row pass: item = (row r, strip of 4 columns); output column x uses window columns x+2 .. x+14
    for (int i = threadIdx.x; i < GA_H * (GA_TW / 4); i += 256) {
        const int r = i >> 4, tx4 = (i & 15) * 4;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float v[20];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(&s[q][r][tx4 + 4 * j]);
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
            float4 o;
            o.x = g13(a.g, v + 2); o.y = g13(a.g, v + 3); o.z = g13(a.g, v + 4); o.w = g13(a.g, v + 5);
            *reinterpret_cast<float4*>(&sr[q][r][tx4]) = o;
        }
    }
    __syncthreads();
    const int tx4 = (threadIdx.x & 15) * 4, y = threadIdx.x >> 4;
    const int gy = y0 + y, gx = x0 + tx4;
    if (gy >= l.h || gx >= l.w) return;
    float b[3][4];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        float c0[13], c1[13], c2[13], c3[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(&sr[q][y + j][tx4]);
            c0[j] = t.x; c1[j] = t.y; c2[j] = t.z; c3[j] = t.w;
        }
        b[q][0] = g13(a.g, c0); b[q][1] = g13(a.g, c1); b[q][2] = g13(a.g, c2); b[q][3] = g13(a.g, c3);
    }
    const size_t o = pb + (size_t)gy * l.pitch + gx;
    const float4 low4 = *reinterpret_cast<const float4*>(a.low + o);
    const float4 rx4 = *reinterpret_cast<const float4*>(a.rx + o);
    const float4 ry4 = *reinterpret_cast<const float4*>(a.ry + o);
    const float lowv[4] = {low4.x, low4.y, low4.z, low4.w}, rxv[4] = {rx4.x, rx4.y, rx4.z, rx4.w}, ryv[4] = {ry4.x, ry4.y, ry4.z, ry4.w};
    float outv[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        // normalize() tail + amplify() (RieszPyramid.cpp:125-143)
        const float tc = __fdiv_rn(b[1][p], b[0][p]), ts = __fdiv_rn(b[2][p], b[0][p]);
        const float mag = __fsqrt_rn(addr(mulr(tc, tc), mulr(ts, ts)));
        float m2 = mulr(mag, a.alpha);
        m2 = (m2 > a.thresh) ? a.thresh : m2;                 // THRESH_TRUNC
        float pc, ps;
        sincosf(m2, &ps, &pc);
        float pair = __fdiv_rn(addr(mulr(rxv[p], tc), mulr(ryv[p], ts)), mag);
        if (pair != pair) pair = 0.f;                          // patchNaNs
        outv[p] = subr(mulr(lowv[p], pc), mulr(pair, ps));
    }
    *reinterpret_cast<float4*>(a.out + o) = make_float4(outv[0], outv[1], outv[2], outv[3]);
}

// result_i = filter2D(injectZerosEven(nearest_up(result_{i+1})), 2*LP) + filter2D(band_i, HP)
// The zero-injected image is never materialised: only the taps that land on even/even fine positions are
// evaluated (25/20/20/16 of 81, by output parity), reading the coarse samples straight from shared memory.
constexpr int RC_CW = R9_SW / 2, RC_CH = R9_SH / 2;   // even fine positions of the window: 36 x 12

template <int PAR, bool MIXED>   // PAR = parity of the output row; MIXED: some of the strip's columns are filter2D tail columns
__device__ __forceinline__ void rc_lowpass(const float (*sc)[RC_CW], int y, int tx, int tail_from, float (&lp)[4]) {
    // output (y, x): taps ky = PAR, PAR+2, ... on coarse row (y + ky) / 2 ; kx likewise by column parity
#pragma unroll
    for (int i = 0; i < 5 - PAR; ++i) {
        const int ky = 2 * i + PAR;
        const float* row = sc[(y + ky) >> 1];
        const float2 q0 = *reinterpret_cast<const float2*>(&row[2 * tx]);
        const float2 q1 = *reinterpret_cast<const float2*>(&row[2 * tx + 2]);
        const float2 q2 = *reinterpret_cast<const float2*>(&row[2 * tx + 4]);
        const float v[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};   // coarse columns x/2 .. x/2+5 (x = 4 tx)
#pragma unroll
        for (int j = 0; j < 5; ++j) {                               // even outputs x, x+2: kx = 0,2,4,6,8
            const float c = 2.0f * c_lp[ky * 9 + 2 * j];
            lp[0] = MIXED ? f2d(c, v[j], lp[0], 0 >= tail_from) : fmaf(c, v[j], lp[0]);
            lp[2] = MIXED ? f2d(c, v[j + 1], lp[2], 2 >= tail_from) : fmaf(c, v[j + 1], lp[2]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {                               // odd outputs x+1, x+3: kx = 1,3,5,7
            const float c = 2.0f * c_lp[ky * 9 + 2 * j + 1];
            lp[1] = MIXED ? f2d(c, v[j + 1], lp[1], 1 >= tail_from) : fmaf(c, v[j + 1], lp[1]);
            lp[3] = MIXED ? f2d(c, v[j + 2], lp[3], 3 >= tail_from) : fmaf(c, v[j + 2], lp[3]);
        }
    }
}

__global__ void __launch_bounds__(256) k_riesz_collapse(Level l, Level lc, const float* __restrict__ band,
                                                        const float* __restrict__ coarse, float* __restrict__ out,
                                                        const __grid_constant__ CUtensorMap tm, int use_tma) {
    __shared__ __align__(128) float sb[R9_SH][R9_SW];
    __shared__ __align__(16) float sc[RC_CH][RC_CW];
    __shared__ __align__(8) uint64_t bar;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * R9_W, y0 = blockIdx.y * R9_H;
    if (use_tma) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
    }
    r9_load_tile(sb, band + (size_t)plane * l.plane, l, x0, y0, use_tma ? &tm : nullptr, plane, &bar);
    const float* __restrict__ c = coarse + (size_t)plane * lc.plane;
    for (int i = threadIdx.x; i < RC_CH * RC_CW; i += 256) {
        const int r = i / RC_CW, cc = i - r * RC_CW;
        // even fine coordinates of the window (reflection keeps parity), value = coarse sample
        const int gy = reflect101(y0 - 4 + 2 * r, l.h), gx = reflect101(x0 - 4 + 2 * cc, l.w);
        sc[r][cc] = __ldg(c + (size_t)(gy >> 1) * lc.pitch + (gx >> 1));
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, y = r9_row(threadIdx.x);
    const int gy = y0 + y, gx = x0 + 4 * tx;
    if (gy >= l.h || gx >= l.w) return;
    float hp[4] = {0.f, 0.f, 0.f, 0.f}, lp[4] = {0.f, 0.f, 0.f, 0.f};
    const int tail_from = (l.w & ~7) - gx;                 // see f2d(): both filter2D calls run at this level's width
    const bool mixed = tail_from < 4;
#pragma unroll
    for (int ky = 0; ky < 9; ++ky) {
        const float4 a0 = *reinterpret_cast<const float4*>(&sb[y + ky][4 * tx]);
        const float4 a1 = *reinterpret_cast<const float4*>(&sb[y + ky][4 * tx + 4]);
        const float4 a2 = *reinterpret_cast<const float4*>(&sb[y + ky][4 * tx + 8]);
        const float v[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
        if (!mixed) {
#pragma unroll
            for (int kx = 0; kx < 9; ++kx) {
                const float cf = c_hp[ky * 9 + kx];
#pragma unroll
                for (int p = 0; p < 4; ++p) hp[p] = fmaf(cf, v[p + kx], hp[p]);
            }
        } else {
#pragma unroll
            for (int kx = 0; kx < 9; ++kx) {
                const float cf = c_hp[ky * 9 + kx];
#pragma unroll
                for (int p = 0; p < 4; ++p) hp[p] = f2d(cf, v[p + kx], hp[p], p >= tail_from);
            }
        }
    }
    if (!mixed) {
        if (y & 1) rc_lowpass<1, false>(sc, y, tx, tail_from, lp);   // warp-uniform (r9_row)
        else rc_lowpass<0, false>(sc, y, tx, tail_from, lp);
    } else {
        if (y & 1) rc_lowpass<1, true>(sc, y, tx, tail_from, lp);
        else rc_lowpass<0, true>(sc, y, tx, tail_from, lp);
    }
    float* o = out + (size_t)plane * l.plane + (size_t)gy * l.pitch + gx;
    *reinterpret_cast<float4*>(o) = make_float4(addr(lp[0], hp[0]), addr(lp[1], hp[1]), addr(lp[2], hp[2]), addr(lp[3], hp[3]));
}

// merge(L', a, b) -> Lab2BGR -> u8 (MagnifyCore.hpp:272-276)
__global__ void __launch_bounds__(256) k_riesz_egress(const float* __restrict__ Lp, Level l, const int16_t* __restrict__ lab,
                                                      int pitch16, size_t plane16, const float4* __restrict__ gtab,
                                                      LabInvCoeffs coeffs, uint8_t* __restrict__ out, size_t step,
                                                      size_t lane_stride, float* __restrict__ fout) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, lane = blockIdx.z;
    if (x >= l.w) return;
    const float L = Lp[(size_t)lane * l.plane + (size_t)y * l.pitch + x];
    const int16_t* p = lab + (size_t)(lane * 3) * plane16 + (size_t)y * pitch16 + x;
    const float A = fmaf((float)p[plane16], 1.0f / 64.0f, -128.0f);
    const float B = fmaf((float)p[2 * plane16], 1.0f / 64.0f, -128.0f);
    float ob, og, orr;
    lab_to_bgr_fast<true>(L, A, B, coeffs, gtab, ob, og, orr);   // NaN L (flat regions) -> white, as OpenCV
    uint8_t* q = out + (size_t)lane * lane_stride + (size_t)y * step + (size_t)x * 3;
    q[0] = unit_to_u8(ob); q[1] = unit_to_u8(og); q[2] = unit_to_u8(orr);
    if (fout) {
        float* f = fout + (((size_t)lane * l.h + y) * l.w + x) * 3;
        f[0] = ob; f[1] = og; f[2] = orr;
    }
}

inline unsigned cdiv(int a, int b) { return (unsigned)((a + b - 1) / b); }

}  // namespace

void RieszMode::reset() {
    arena.release();
    lv.clear();
    for (auto* v : {&oct, &cur_low, &cur_rx, &cur_ry, &old_low, &old_rx, &old_ry, &phase_c, &phase_s, &lo_r0c, &lo_r0s,
                    &lo_r1c, &lo_r1s, &hi_r0c, &hi_r0s, &hi_r1c, &hi_r1s, &amp, &t_c, &t_s, &low_amp, &res})
        v->clear();
    lab16 = nullptr;
    allocated = false;
}

#define KLAUNCH(name, level, ...)                                                  \
    do {                                                                           \
        const bool p__ = ctx.prof && ctx.prof->begin(name, level, ctx.stream);     \
        __VA_ARGS__;                                                               \
        if (p__) ctx.prof->end(ctx.stream);                                        \
        MCK(cudaGetLastError());                                                   \
        ++*ctx.launches;                                                           \
    } while (0)

mc_status RieszMode::build_pyramid(const ModeCtx& ctx) {
    // RieszPyramid::buildPyramid (RieszPyramid.cpp:215-238); the Riesz pair itself is formed in the phase kernel
    for (int i = 0; i < levels - 1; ++i) {
        const Level& l = lv[(size_t)i];
        dim3 grid(cdiv(l.w, R9_W), cdiv(l.h, R9_H), lanes);
        const int tma = ctx.use_tma && tm_valid[(size_t)i];
        KLAUNCH("riesz_analysis", i, k_riesz_analysis<<<grid, 256, 0, ctx.stream>>>(l, lv[(size_t)i + 1], oct[(size_t)i], cur_low[(size_t)i], oct[(size_t)i + 1],
                                                                                    *reinterpret_cast<const CUtensorMap*>(&tm_oct[(size_t)i]), tma));
    }
    return MC_OK;
}

mc_status RieszMode::process(const ModeCtx& ctx, const FrameIO& io, const mc_params& p, int nlevels, int* produced) {
    *produced = 0;
    if (io.channels < 3) return MC_OK;  // MagnifyCore.hpp:212: gray input is a silent passthrough
    const bool first = !allocated || std::isnan(loA[0]) || std::isnan(hiA[0]);  // :226
    if (first) {
        reset();
        levels = nlevels; w = io.w; h = io.h;
        lv.resize((size_t)levels);
        int cw = w, ch = h;
        for (int i = 0; i < levels; ++i) {
            lv[(size_t)i] = make_level(cw, ch);
            cw = cw / 2 + (cw % 2); ch = ch / 2 + (ch % 2);   // subsample(), RieszPyramid.cpp:262-263
        }
        auto alloc_set = [&](std::vector<float*>& v, int n_levels, bool zero) -> mc_status {
            v.assign((size_t)levels, nullptr);
            for (int i = 0; i < n_levels; ++i) {
                const size_t n = (size_t)lanes * lv[(size_t)i].plane;
                MCK(arena.alloc(&v[(size_t)i], n));
                if (zero) MCK(cudaMemsetAsync(v[(size_t)i], 0, n * sizeof(float), ctx.stream));
            }
            return MC_OK;
        };
        mc_status st;
        const int nb = levels - 1;  // band levels
        if ((st = alloc_set(oct, levels, false)) != MC_OK) return st;
        if ((st = alloc_set(cur_low, nb, false)) != MC_OK) return st;
        for (auto* v : {&cur_rx, &cur_ry, &old_low, &amp, &t_c, &t_s, &low_amp})
            if ((st = alloc_set(*v, nb, false)) != MC_OK) return st;
        if ((st = alloc_set(res, levels, false)) != MC_OK) return st;
        // init(): Riesz pair of `old`, phases and IIR registers start at zero (RieszPyramid.cpp:196-213, TemporalFilter.cpp:299-317)
        for (auto* v : {&old_rx, &old_ry, &phase_c, &phase_s, &lo_r0c, &lo_r0s, &lo_r1c, &lo_r1s, &hi_r0c, &hi_r0s, &hi_r1c, &hi_r1s})
            if ((st = alloc_set(*v, nb, true)) != MC_OK) return st;
        // TMA descriptors for the tiles of the 9x9 kernels (option use_tma; tiles touching a border keep the reflected loads)
        tm_oct.assign((size_t)levels, TensorMapStorage{});
        tm_band.assign((size_t)levels, TensorMapStorage{});
        tm_valid.assign((size_t)levels, 0);
        for (int i = 0; i < nb; ++i)
            tm_valid[(size_t)i] = make_tensor_map_box(&tm_oct[(size_t)i], oct[(size_t)i], lv[(size_t)i], lanes, R9_SW, R9_SH) &&
                                  make_tensor_map_box(&tm_band[(size_t)i], low_amp[(size_t)i], lv[(size_t)i], lanes, R9_SW, R9_SH) ? 1 : 0;
        pitch16 = round_up(w, 64);
        plane16 = (size_t)h * pitch16;
        void* lp = nullptr;
        MCK(arena.alloc_bytes(&lp, (size_t)lanes * 3 * plane16 * sizeof(int16_t)));
        lab16 = (int16_t*)lp;
        lo_freq = p.coLow; hi_freq = p.coHigh; framerate = p.framerate;
        auto design = [&](double f, double* A, double* B) {
            const double wn = framerate == 0.0 ? 0.0 : f / (framerate / 2.0);  // TemporalFilter.cpp:324-327
            std::vector<double> a, b;
            butterworth(2, wn, a, b);
            for (int k = 0; k < 3; ++k) { A[k] = a[(size_t)k]; B[k] = b[(size_t)k]; }
        };
        design(lo_freq, loA, loB);
        design(hi_freq, hiA, hiB);
        allocated = true;
    }
    // BGR -> Lab; only L is magnified (MagnifyCore.hpp:217-222)
    LAUNCH("lab16", 0, launch_lab16(io, *ctx.tables, lab16, pitch16, plane16, ctx.stream, oct[0], lv[0].pitch, lv[0].plane));
    mc_status st = build_pyramid(ctx);
    if (st != MC_OK) return st;
    if (first) {
        // old = pyramid of the first frame with a zero Riesz pair; the frame itself is shown unmagnified (:239)
        for (int i = 0; i < levels - 1; ++i) std::swap(cur_low[(size_t)i], old_low[(size_t)i]);
        return MC_OK;
    }
    // cutoff changes re-design the filter, zero both filters' registers and rebuild `old` from this frame (:243-254)
    bool rebuild_old = false;
    auto zero_state = [&]() -> mc_status {
        for (auto* v : {&phase_c, &phase_s, &lo_r0c, &lo_r0s, &lo_r1c, &lo_r1s, &hi_r0c, &hi_r0s, &hi_r1c, &hi_r1s})
            for (int i = 0; i < levels - 1; ++i)
                MCK(cudaMemsetAsync((*v)[(size_t)i], 0, (size_t)lanes * lv[(size_t)i].plane * sizeof(float), ctx.stream));
        return MC_OK;
    };
    auto redesign = [&](double f, double* A, double* B) {
        const double wn = framerate == 0.0 ? 0.0 : f / (framerate / 2.0);
        std::vector<double> a, b;
        butterworth(2, wn, a, b);
        for (int k = 0; k < 3; ++k) { A[k] = a[(size_t)k]; B[k] = b[(size_t)k]; }
    };
    if (lo_freq != p.coLow) {
        lo_freq = p.coLow;
        redesign(lo_freq, loA, loB);
        if ((st = zero_state()) != MC_OK) return st;
        rebuild_old = true;
    }
    if (hi_freq != p.coHigh) {
        hi_freq = p.coHigh;
        redesign(hi_freq, hiA, hiB);
        if ((st = zero_state()) != MC_OK) return st;
        rebuild_old = true;
    }
    const int nb = levels - 1;
    for (int i = 0; i < nb; ++i) {
        PhaseArgs a;
        a.l = lv[(size_t)i];
        a.low = cur_low[(size_t)i];
        a.plow = rebuild_old ? nullptr : old_low[(size_t)i];
        a.prx = old_rx[(size_t)i]; a.pry = old_ry[(size_t)i];
        a.rx = cur_rx[(size_t)i]; a.ry = cur_ry[(size_t)i];
        a.ph_c = phase_c[(size_t)i]; a.ph_s = phase_s[(size_t)i];
        a.lo_r0c = lo_r0c[(size_t)i]; a.lo_r0s = lo_r0s[(size_t)i]; a.lo_r1c = lo_r1c[(size_t)i]; a.lo_r1s = lo_r1s[(size_t)i];
        a.hi_r0c = hi_r0c[(size_t)i]; a.hi_r0s = hi_r0s[(size_t)i]; a.hi_r1c = hi_r1c[(size_t)i]; a.hi_r1s = hi_r1s[(size_t)i];
        a.amp = amp[(size_t)i]; a.t_c = t_c[(size_t)i]; a.t_s = t_s[(size_t)i];
        a.lo = Butter{loB[0], loB[1], loB[2], loA[1], loA[2]};
        a.hi = Butter{hiB[0], hiB[1], hiB[2], hiA[1], hiA[2]};
        dim3 grid(cdiv(a.l.w, RT_W), cdiv(a.l.h, RT_H), lanes);
        KLAUNCH("riesz_phase", i, k_riesz_phase<<<grid, 256, 0, ctx.stream>>>(a));
    }
    // *old = *cur (before amplification, :267): the buffers just written become `old`
    for (int i = 0; i < nb; ++i) {
        std::swap(cur_low[(size_t)i], old_low[(size_t)i]);
        std::swap(cur_rx[(size_t)i], old_rx[(size_t)i]);
        std::swap(cur_ry[(size_t)i], old_ry[(size_t)i]);
    }
    if (ctx.analysis_only) {   // state-carry pass of temporal sharding: pyramids and filter registers are up to date
        *produced = 0;
        return MC_OK;
    }
    // amplify (RieszPyramid.cpp:248-252) — this frame's band/pair now live in the old_* buffers
    const float alpha = (float)p.amplification;
    const float thresh = (float)(p.coWavelength * (3.14159265358979323846 / 100.0));  // PI_PERCENT, :214,:269
    for (int i = nb - 1; i >= 0; --i) {
        AmpArgs a;
        a.l = lv[(size_t)i];
        a.amp = amp[(size_t)i]; a.t_c = t_c[(size_t)i]; a.t_s = t_s[(size_t)i];
        a.low = old_low[(size_t)i]; a.rx = old_rx[(size_t)i]; a.ry = old_ry[(size_t)i];
        a.out = low_amp[(size_t)i];
        gaussian_kernel_13_3(a.g.k);
        a.alpha = alpha; a.thresh = thresh;
        dim3 grid(cdiv(a.l.w, GA_TW), cdiv(a.l.h, GA_TH), lanes);
        KLAUNCH("riesz_amplify", i, k_riesz_amplify<<<grid, 256, 0, ctx.stream>>>(a));
    }
    // collapse (RieszPyramid.cpp:304-325)
    const float* result = oct[(size_t)levels - 1];
    for (int i = nb - 1; i >= 0; --i) {
        const Level& l = lv[(size_t)i];
        dim3 grid(cdiv(l.w, R9_W), cdiv(l.h, R9_H), lanes);
        const int tma = ctx.use_tma && tm_valid[(size_t)i];
        KLAUNCH("riesz_collapse", i, k_riesz_collapse<<<grid, 256, 0, ctx.stream>>>(l, lv[(size_t)i + 1], low_amp[(size_t)i], result, res[(size_t)i],
                                                                                    *reinterpret_cast<const CUtensorMap*>(&tm_band[(size_t)i]), tma));
        result = res[(size_t)i];
    }
    {
        dim3 grid(cdiv(w, 256), h, lanes);
        KLAUNCH("riesz_egress", 0, k_riesz_egress<<<grid, 256, 0, ctx.stream>>>(result, lv[0], lab16, pitch16, plane16, ctx.tables->inv_gamma, ctx.tables->inv_coeffs, io.out, io.out_step, io.out_lane_stride, ctx.float_out));
    }
    *produced = 1;
    return MC_OK;
}

void RieszMode::find_state(const char* name, int level, StateRef& out) {
    out = StateRef{};
    if (!allocated || level < 0 || level >= levels - 1) return;
    struct { const char* n; std::vector<float*>* v; } map[] = {
        {"old.lowpass", &old_low}, {"old.rx", &old_rx}, {"old.ry", &old_ry}, {"phase.c", &phase_c}, {"phase.s", &phase_s},
        {"lo.r0.c", &lo_r0c}, {"lo.r0.s", &lo_r0s}, {"lo.r1.c", &lo_r1c}, {"lo.r1.s", &lo_r1s},
        {"hi.r0.c", &hi_r0c}, {"hi.r0.s", &hi_r0s}, {"hi.r1.c", &hi_r1c}, {"hi.r1.s", &hi_r1s}};
    for (auto& m : map)
        if (!std::strcmp(name, m.n)) {
            const Level& l = lv[(size_t)level];
            out.ptr = (*m.v)[(size_t)level]; out.rows = l.h; out.cols = l.w; out.channels = 1; out.pitch = l.pitch; out.plane_stride = l.plane;
            return;
        }
}

}  // namespace mc
