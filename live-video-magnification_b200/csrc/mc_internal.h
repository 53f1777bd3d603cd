// Internal interfaces between the handle (mc_core.cu), the tables (mc_tables.cpp) and the kernel
// launchers (mc_laplace.cu, mc_color.cu, mc_riesz.cu).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "mc_math.cuh"

namespace mc {

// Geometry of one pyramid level; planes are f32 [plane][h][pitch] with pitch % 32 == 0.
struct Level {
    int w = 0, h = 0, pitch = 0;
    size_t plane = 0;  // floats per plane = h * pitch
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

inline Level make_level(int w, int h) {
    Level l;
    l.w = w; l.h = h; l.pitch = round_up(w, 32); l.plane = (size_t)l.h * l.pitch;
    return l;
}

// Device constant tables shared by every handle on a device (built once per device).
struct DeviceTables {
    LabLutEntry* lab_lut = nullptr;   // [33][33][33]
    float4* inv_gamma = nullptr;      // [1024] spline coefficients {f, b, c, d}
    LabInvCoeffs inv_coeffs{};
};

// mc_tables.cpp ------------------------------------------------------------------------------
void build_lab_lut_packed(std::vector<LabLutEntry>& out);          // from the embedded int16 table
void build_inv_gamma_spline(std::vector<float4>& out);             // OpenCV sRGBInvGammaTab
void build_lab_inv_coeffs(LabInvCoeffs& out);
int calculate_max_levels(int w, int h);
int optimal_buffer_size(int fps);
void butterworth(unsigned order, double wn, std::vector<double>& a, std::vector<double>& b);
void motion_gains(double amplification, double coWavelength, int levels, int w, int h, std::vector<float>& gains);
void gaussian_kernel_13_3(float taps[13]);

// ------------------------------------------------------------------------------------------------
// Laplace launchers (mc_laplace.cu).  All take the handle's stream; every call is one kernel launch
// and returns the launch's cudaError_t (cudaGetLastError()).
// ------------------------------------------------------------------------------------------------
struct FrameIO {
    const uint8_t* in = nullptr;   // [lanes][h][in_step]
    size_t in_step = 0, in_lane_stride = 0;
    uint8_t* out = nullptr;
    size_t out_step = 0, out_lane_stride = 0;
    int w = 0, h = 0, channels = 0, lanes = 0;
};

// u8 frame -> (Lab | gray) -> pyrDown -> G1 planes.  Optionally materialises G0 planes (faithful mode).
cudaError_t launch_ingest_down(const FrameIO& io, const DeviceTables& tb, const Level& l0, const Level& l1,
                               float* g1, float* g0_or_null, cudaStream_t s);

struct LevelArgs {
    Level lf, lc;              // this level (fine) and the next (coarse)
    const float* g;            // G_l planes (fine)
    float* g_next;             // G_{l+1} planes (written)
    float* hi; float* lo;      // state planes of this level
    float* m;                  // gain * (hi - lo), may be null
    int planes;
    int first;                 // 1: hi = lo = band (MagnifyCore.hpp:98-103)
    double c_hi, one_minus_c_hi, c_lo, one_minus_c_lo;
    float gain;
};
// fused per level: pyrDown + pyrUp + subtract + dual-EMA update + gain (SpatialFilter.cpp:25-38,
// TemporalFilter.cpp:9-22, MagnifyCore.hpp:127-134)
cudaError_t launch_level(const LevelArgs& a, cudaStream_t s);

// cur_l = pyrUp(cur_{l+1}) + m_l, in place in m_l (SpatialFilter.cpp:52-61)
cudaError_t launch_collapse(const Level& lf, const Level& lc, float* m_fine, const float* m_coarse, int planes,
                            cudaStream_t s);

// out = convert(input + chroma * pyrUp(cur_1)) (MagnifyCore.hpp:136-158).  m1 == nullptr: no motion.
cudaError_t launch_egress(const FrameIO& io, const DeviceTables& tb, const Level& l0, const Level& l1,
                          const float* m1, float chroma, float* float_out_or_null, cudaStream_t s);

// plane copy helpers
cudaError_t launch_copy_planes(float* dst, const float* src, size_t n, cudaStream_t s);

}  // namespace mc
