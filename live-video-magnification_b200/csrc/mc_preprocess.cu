// Front of the reference chain fused onto the device: PreprocessProcessor (normalised ROI crop +
// INTER_AREA 1/2..1/8 downscale, reference src/processing/PreprocessProcessor.cpp:10-51) and GrayscaleProcessor
// (BGR2GRAY, src/processing/GrayscaleProcessor.cpp:7-16), bit-exact with OpenCV's u8 paths (mc_math.cuh).
#include "mc_internal.h"

namespace mc {

namespace {

struct PreArgs {
    const uint8_t* src;   // ROI origin inside the raw frame
    size_t step;
    int cn, dw, dh, isx, isy, fast, copy_only;
    const AreaTap* xtab; const int* xofs; const AreaTap* ytab; const int* yofs;
    uint8_t* dst;         // [dh][dw][cn] tight, or null
    uint8_t* gray;        // [dh][dw] tight, or null (cn == 3 only)
};

__global__ void __launch_bounds__(256) k_preprocess(const PreArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.dw) return;
    int v[3] = {0, 0, 0};
    for (int ch = 0; ch < a.cn; ++ch) {
        v[ch] = a.copy_only ? (int)__ldg(a.src + (size_t)y * a.step + (size_t)x * a.cn + ch)
                            : (int)resize_area_sample(a.src, a.step, a.cn, ch, y, x, a.isx, a.isy, a.fast != 0, a.xtab, a.xofs, a.ytab, a.yofs);
        if (a.dst) a.dst[((size_t)y * a.dw + x) * a.cn + ch] = (uint8_t)v[ch];
    }
    if (a.gray) a.gray[(size_t)y * a.dw + x] = bgr_to_gray_u8(v[0], v[1], v[2]);
}

}  // namespace

cudaError_t launch_preprocess(const uint8_t* src_roi, size_t step, int cn, int sw, int sh, int dw, int dh, bool copy_only,
                              const AreaTap* xtab, const int* xofs, const AreaTap* ytab, const int* yofs, uint8_t* dst,
                              uint8_t* gray, cudaStream_t s) {
    PreArgs a;
    a.src = src_roi; a.step = step; a.cn = cn; a.dw = dw; a.dh = dh; a.copy_only = copy_only ? 1 : 0;
    const double sx = (double)sw / dw, sy = (double)sh / dh;
    a.isx = (int)sx; a.isy = (int)sy;
    a.fast = (std::abs(sx - a.isx) < 2.220446049250313e-16 && std::abs(sy - a.isy) < 2.220446049250313e-16) ? 1 : 0;  // DBL_EPSILON
    a.xtab = xtab; a.xofs = xofs; a.ytab = ytab; a.yofs = yofs; a.dst = dst; a.gray = gray;
    dim3 grid((unsigned)((dw + 255) / 256), (unsigned)dh, 1);
    k_preprocess<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace mc
