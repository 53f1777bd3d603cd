// Per-mode device state + per-frame drivers: the B200 twins of magcore::MotionState / ColorState /
// RieszState and magnifyMotion / magnifyColor / magnifyRiesz (reference
// src/processing/magnification/MagnifyCore.hpp:24-40, :83-279).
#pragma once
#include <cufft.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/magcore_b200.h"
#include "mc_internal.h"

namespace mc {

// Optional per-launch event timing (bench roofline); events are drained by mc_profile_read.
struct Profiler {
    struct Rec { const char* name; int level; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    bool begin(const char* name, int level, cudaStream_t s) {
        Rec r{name, level, nullptr, nullptr};
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return false;
        cudaEventRecord(r.a, s);
        recs.push_back(r);
        return true;
    }
    void end(cudaStream_t s) { cudaEventRecord(recs.back().b, s); }
};

struct ModeCtx {
    cudaStream_t stream;
    const DeviceTables* tables;
    uint64_t* launches;
    std::string* err;
    bool faithful0;
    float* float_out;  // optional [lanes][h][w][C] pre-quantisation tap
    Profiler* prof;    // optional
    bool use_tma;      // stage level-kernel tiles with cp.async.bulk.tensor (option "use_tma", default on)
    bool prefetch_state;    // level kernel requests its state tiles by TMA at kernel entry (option "prefetch_state", default on:
                            // B200, 32 lanes: level[1] 235 -> 205 us)
    int egress_strip;       // Laplace egress as the register/shuffle strip kernel (option "egress_strip": 0 = tile kernel,
                            // 16 / 20 / 24 = strip kernel compiled for that many resident warps per SM)
    int ingest_warps;       // warps per CTA of the fused ingest kernel (option "ingest_warps": 1, 2 or 4)
    bool band_from_state;   // synthesis rebuilds gain*(hi-lo) from the state planes instead of reading a stored band
                            // (option "band_from_state", default on: with prefetch_state level[1] 205 -> 177 us, egress +8 us)
    int lane_groups;        // Laplace: number of lane groups run as concurrent launch chains (option "lane_groups"; 0 = automatic)
    bool analysis_only;     // Laplace / Phase: update the temporal state but skip synthesis + egress (*produced = 0); used by the
                            // state-carry pass of temporal sharding (lvm_b200.shard.magnify_segment)
};

// Launch bookkeeping shared by the mode drivers: counts the launch, optionally brackets it with events.
#define MCK(call)                                                             \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) {                                             \
            *ctx.err = std::string(#call) + ": " + cudaGetErrorString(e__);   \
            return MC_ERR_CUDA;                                               \
        }                                                                     \
    } while (0)
#define MCK_ST(call)                                                          \
    do {                                                                      \
        const mc_status s__ = (call);                                         \
        if (s__ != MC_OK) return s__;                                         \
    } while (0)
#define LAUNCH(name, level, call)                                             \
    do {                                                                      \
        const bool p__ = ctx.prof && ctx.prof->begin(name, level, ctx.stream);\
        cudaError_t e__ = (call);                                             \
        if (p__) ctx.prof->end(ctx.stream);                                   \
        if (e__ != cudaSuccess) {                                             \
            *ctx.err = std::string(name) + ": " + cudaGetErrorString(e__);    \
            return MC_ERR_CUDA;                                               \
        }                                                                     \
        ++*ctx.launches;                                                      \
    } while (0)

struct StateRef {  // a named state plane set for mc_get_state / mc_set_state
    float* ptr = nullptr;
    int rows = 0, cols = 0, channels = 0, pitch = 0;
    size_t plane_stride = 0;
};

// Simple owner of cudaMalloc'ed float buffers (freed together on reset()).
struct DeviceArena {
    std::vector<void*> blocks;
    cudaError_t alloc(float** p, size_t floats);
    cudaError_t alloc_bytes(void** p, size_t bytes);
    void free_block(void* p);   // returns one block early (buffers replaced when a ring grows)
    void release();
};

struct MotionMode {
    int lanes = 1;
    bool empty = true;       // MotionState::empty()
    bool allocated = false;
    int levels = 0, channels = 0, w = 0, h = 0;
    bool faithful = false;
    bool from_state = false;   // ModeCtx::band_from_state at allocation time
    std::vector<Level> lv;              // 0..levels
    std::vector<float*> G, hi, lo, M;   // per level (null where not kept)
    int16_t* lab16 = nullptr;           // Lab planes of the current frame (C == 3)
    int pitch16 = 0;
    size_t plane16 = 0;
    DeviceArena arena;

    // Lane groups (option "lane_groups"): the streams of a handle are independent, so their launch sets are issued as
    // `groups.size()` separate chains on separate CUDA streams.  A group's kernels no longer fill the GPU, so the block
    // scheduler co-schedules different stages of different groups on the same SMs — the L1-bound BGR->Lab ingest of one
    // group, the issue-bound egress of another and the HBM-bound level kernels of a third — instead of running the
    // stages back to back.  Buffers stay whole-handle allocations; a group sees them through offset pointers.
    struct Group {
        int lane0 = 0, lanes = 0;
        cudaStream_t stream = nullptr;     // null: the handle's stream (single group)
        cudaEvent_t done = nullptr;
        std::vector<TensorMapStorage> tmaps, tmaps_hi, tmaps_lo;   // per level: TMA descriptors of G[l] / hi[l] / lo[l] of this group's planes
        std::vector<char> tmap_valid;
    };
    std::vector<Group> groups;
    cudaEvent_t ev_fork = nullptr;
    int groups_req = 0;                    // ModeCtx::lane_groups at allocation time
    std::vector<float> gains;              // per-level gains of the current frame (member: no per-frame allocation)

    void reset();
    mc_status process(const ModeCtx& ctx, const FrameIO& io, const mc_params& p, int levels, int* produced);
    void find_state(const char* name, int level, StateRef& out);

private:
    mc_status allocate(const ModeCtx& ctx, const FrameIO& io, int levels);
    mc_status make_groups(const ModeCtx& ctx);
    void drop_groups();
    mc_status run_group(const ModeCtx& ctx, const FrameIO& io, const mc_params& p, Group& g, bool first, double c_lo, double c_hi);
};

struct ColorMode {
    int lanes = 1;
    bool allocated = false;
    int levels = 0, channels = 0, w = 0, h = 0;
    int count = 0;          // frames currently in the window
    int head = 0;           // physical slot of the OLDEST column
    std::vector<Level> lv;  // gaussian chain levels 0..levels
    std::vector<float*> G;  // pyrDown chain scratch (levels 1..levels-1) ; small level goes to the ring
    std::vector<float*> U;  // up-chain scratch
    std::vector<Level> ulv;
    float* ring = nullptr;      // [planes][rowsP][cap] time-major rows: each pixel's samples contiguous
    float* work = nullptr;      // gathered/filtered window [planes*rows][n]
    cufftComplex* spec = nullptr;
    float* minmax = nullptr;    // device scalars
    int small_rows = 0;         // pixels of the small level
    int ring_cap = 0;           // physical slots allocated
    int ring_mod = 0;           // logical ring size (<= ring_cap): max(getOptimalBufferSize(int(framerate)), window length)
    struct FftPlans { cufftHandle r2c = 0, c2r = 0; size_t work_bytes = 0; void* bound = nullptr; };
    std::map<int, FftPlans> plans;   // per DFT length (the window length: 2 ... cap during warm-up), for plan_signals signals
    size_t plan_signals = 0;
    void* fft_work = nullptr;        // one work area shared by all cached plans (they run back to back on one stream)
    size_t fft_work_bytes = 0;
    DeviceArena arena;

    void reset();
    mc_status process(const ModeCtx& ctx, const FrameIO& io, const mc_params& p, int levels, int* produced);
    void find_state(const char* name, int level, StateRef& out);
};

struct RieszMode {
    int lanes = 1;
    bool allocated = false;     // st.cur != null
    int levels = 0, w = 0, h = 0;
    double lo_freq = 0, hi_freq = 0, framerate = 0;   // itsFrequency / itsFramerate of the two filters
    double loA[3] = {1, 0, 0}, loB[3] = {0, 0, 0}, hiA[3] = {1, 0, 0}, hiB[3] = {0, 0, 0};
    std::vector<Level> lv;      // octave sizes, 0..levels-1 (band levels 0..levels-2 + low-pass residual)
    DeviceArena arena;
    // per level planes (one plane per lane: only L is processed)
    std::vector<float*> oct;                                  // octave i (input of level i); oct[levels-1] is the residual
    std::vector<float*> cur_low, cur_rx, cur_ry;              // this frame's band + Riesz pair
    std::vector<float*> old_low, old_rx, old_ry;              // prior pyramid (RieszState::old)
    std::vector<float*> phase_c, phase_s;                     // accumulated phase (the two filters' copies are identical)
    std::vector<float*> lo_r0c, lo_r0s, lo_r1c, lo_r1s, hi_r0c, hi_r0s, hi_r1c, hi_r1s;   // DF-II registers
    std::vector<float*> amp, t_c, t_s, low_amp, res;
    int16_t* lab16 = nullptr;
    int pitch16 = 0;
    size_t plane16 = 0;
    // TMA descriptors of the 9x9 kernels' input tiles (72 x 24 boxes): octave i (analysis), amplified band i (collapse)
    std::vector<TensorMapStorage> tm_oct, tm_band;
    std::vector<char> tm_valid;

    void reset();
    mc_status process(const ModeCtx& ctx, const FrameIO& io, const mc_params& p, int levels, int* produced);
    void find_state(const char* name, int level, StateRef& out);

private:
    mc_status build_pyramid(const ModeCtx& ctx);
};

}  // namespace mc
