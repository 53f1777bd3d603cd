// The handle behind the C ABI: a B200-resident twin of livim::MagnificationProcessor
// (reference src/processing/MagnificationProcessor.cpp:10-67) — level clamp, structural reset,
// mode dispatch, passthrough decisions — plus device memory, streams and the pinned pipeline.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "../../include/magcore_b200.h"
#include "mc_internal.h"
#include "mc_modes.h"

using namespace mc;

namespace {
thread_local std::string g_create_error;

struct Slot {  // one in-flight frame of the pinned pipeline
    uint8_t *h_in = nullptr, *h_out = nullptr;   // pinned staging (lanes frames)
    uint8_t *d_in = nullptr, *d_out = nullptr;   // device frames
    size_t bytes = 0;
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_done = nullptr;
    int produced = 0;
    uint8_t* user_out = nullptr;                 // destination given to mc_submit
    size_t user_out_step = 0;
    bool direct_out = false;                     // D2H went straight into user_out (pinned)
    int w = 0, h = 0, c = 0;
};
}  // namespace

struct mc_handle {
    int device = 0;
    int lanes = 1;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    DeviceTables tables;
    std::string err;
    uint64_t launches = 0;

    // StructuralTracker (MagnifyCore.hpp:45-80)
    int t_mode = MC_MODE_NONE, t_levels = -1, t_channels = -1, t_w = 0, t_h = 0;
    int t_down = 1, t_roi = 0;
    float t_rx = 0.f, t_ry = 0.f, t_rw = 1.f, t_rh = 1.f;

    // options
    bool faithful0 = false, keep_float = false, profile = false, use_tma = true, prefetch_state = true, band_from_state = true, analysis_only = false;
    int ingest_warps = 1;
    int egress_strip = 20;
    int lane_groups = 0;
    Profiler prof;
    int depth = 3;

    MotionMode motion;
    ColorMode color;
    RieszMode riesz;

    float* float_out = nullptr;
    size_t float_out_floats = 0;

    // mc_chain_process scratch (raw frame, preprocessed frame, gray frame, magnified frame, INTER_AREA taps)
    uint8_t *c_raw = nullptr, *c_pre = nullptr, *c_gray = nullptr, *c_out = nullptr;
    size_t c_raw_b = 0, c_pre_b = 0, c_gray_b = 0, c_out_b = 0;
    void* c_tabs = nullptr;
    size_t c_tabs_b = 0;

    // pipeline
    std::vector<Slot> slots;
    std::deque<int> inflight;
    int next_slot = 0;
};

namespace {
// ---- exception firewall of the C ABI ------------------------------------------------------------------------
// include/magcore_b200.h promises that no C++ exception crosses the boundary (the reference's own firewall,
// ProcessingChain.cpp:50-62, sits ABOVE the adapter and only understands what the adapter rethrows).  Every
// extern "C" entry with a body that can allocate is a function-try-block ending in on_exception().
thread_local int g_inject_throw = 0;   // test hook (mc_debug_inject_exception): >0 = throw at the n-th guarded entry

inline void debug_maybe_throw() {
    if (g_inject_throw > 0 && --g_inject_throw == 0) throw std::bad_alloc();
}

void reset_modes(mc_handle* h);
void tracker_reset(mc_handle* h);

mc_status on_exception(mc_handle* h) noexcept {
    const char* what = "unknown C++ exception";
    std::string buf;
    try { throw; }
    catch (const std::bad_alloc&) { what = "out of host memory (std::bad_alloc)"; }
    catch (const std::exception& e) {
        try { buf = std::string("C++ exception: ") + e.what(); what = buf.c_str(); } catch (...) {}
    }
    catch (...) {}
    try {
        if (h) {
            h->err = what;
            // the recovery contract of ProcessingChain.cpp:50-62: temporal state may be half-updated, drop it
            reset_modes(h);
            tracker_reset(h);
        } else {
            g_create_error = what;
        }
    } catch (...) {}
    return MC_ERR_INTERNAL;
}
}  // namespace

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                         \
            return MC_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

namespace {

void tracker_disable(mc_handle* h) {
    h->t_mode = MC_MODE_NONE; h->t_levels = -1; h->t_channels = -1; h->t_w = 0; h->t_h = 0;
}
void tracker_reset(mc_handle* h) {
    tracker_disable(h);
    h->t_down = 1; h->t_roi = 0; h->t_rx = 0.f; h->t_ry = 0.f; h->t_rw = 1.f; h->t_rh = 1.f;
}
bool tracker_update(mc_handle* h, const mc_params* p, int lv, int ch, int w, int hh) {
    const bool change = p->mode != h->t_mode || lv != h->t_levels || w != h->t_w || hh != h->t_h ||
                        ch != h->t_channels || p->pre_downscale != h->t_down ||
                        (p->pre_roiEnabled != 0) != (h->t_roi != 0) || p->pre_roiX != h->t_rx ||
                        p->pre_roiY != h->t_ry || p->pre_roiW != h->t_rw || p->pre_roiH != h->t_rh;
    if (change) {
        h->t_mode = p->mode; h->t_levels = lv; h->t_w = w; h->t_h = hh; h->t_channels = ch;
        h->t_down = p->pre_downscale; h->t_roi = p->pre_roiEnabled != 0;
        h->t_rx = p->pre_roiX; h->t_ry = p->pre_roiY; h->t_rw = p->pre_roiW; h->t_rh = p->pre_roiH;
    }
    return change;
}

void reset_modes(mc_handle* h) {
    h->motion.reset();
    h->color.reset();
    h->riesz.reset();
}

void free_slots(mc_handle* h) {
    for (auto& s : h->slots) {
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_k) cudaEventDestroy(s.ev_k);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
    }
    h->slots.clear();
    h->inflight.clear();
    h->next_slot = 0;
}

mc_status ensure_slots(mc_handle* h, size_t bytes) {
    if ((int)h->slots.size() == h->depth && !h->slots.empty() && h->slots[0].bytes >= bytes) return MC_OK;
    if (!h->inflight.empty()) {
        h->err = "pipeline geometry changed with frames in flight";
        return MC_ERR_INVALID;
    }
    free_slots(h);
    h->slots.resize((size_t)h->depth);
    for (auto& s : h->slots) {
        s.bytes = bytes;
        CK(cudaHostAlloc((void**)&s.h_in, bytes, cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&s.h_out, bytes, cudaHostAllocDefault));
        CK(cudaMalloc((void**)&s.d_in, bytes));
        CK(cudaMalloc((void**)&s.d_out, bytes));
        CK(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    }
    return MC_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// The body of MagnificationProcessor::process on device-resident frames.
mc_status process_device_impl(mc_handle* h, const uint8_t* d_in, int w, int hh, int channels, size_t in_step,
                              const mc_params* p, uint8_t* d_out, size_t out_step, int* produced) {
    *produced = 0;
    debug_maybe_throw();
    if (!p) { h->err = "params is null"; return MC_ERR_INVALID; }
    // Identity when disabled / empty; free state so a later re-enable starts cleanly (:21-29).
    if (p->mode == MC_MODE_NONE || d_in == nullptr || w <= 0 || hh <= 0) {
        if (h->t_mode != MC_MODE_NONE) {
            reset_modes(h);
            tracker_disable(h);
        }
        return MC_OK;
    }
    if (p->mode < 0 || p->mode > MC_MODE_NONE) { h->err = "bad mode"; return MC_ERR_INVALID; }
    if (channels != 1 && channels != 3) { h->err = "channels must be 1 or 3"; return MC_ERR_INVALID; }
    if (d_out == nullptr) { h->err = "output pointer is null"; return MC_ERR_INVALID; }
    if (in_step < (size_t)w * channels || out_step < (size_t)w * channels) { h->err = "step too small"; return MC_ERR_INVALID; }
    const int max_levels = calculate_max_levels(w, hh);  // :32-33
    if (max_levels < 1) return MC_OK;
    const int levels = std::min(std::max((int)p->levels, 1), max_levels);  // :34
    if (tracker_update(h, p, levels, channels, w, hh)) reset_modes(h);      // :39-43

    FrameIO io;
    io.in = d_in; io.in_step = in_step; io.in_lane_stride = in_step * (size_t)hh;
    io.out = d_out; io.out_step = out_step; io.out_lane_stride = out_step * (size_t)hh;
    io.w = w; io.h = hh; io.channels = channels; io.lanes = h->lanes;

    float* fout = nullptr;
    if (h->keep_float) {
        const size_t need = (size_t)h->lanes * w * hh * channels;
        if (need > h->float_out_floats) {
            if (h->float_out) cudaFree(h->float_out);
            h->float_out = nullptr;
            CK(cudaMalloc((void**)&h->float_out, need * sizeof(float)));
            h->float_out_floats = need;
        }
        fout = h->float_out;
    }

    ModeCtx ctx{h->stream, &h->tables, &h->launches, &h->err, h->faithful0, fout, h->profile ? &h->prof : nullptr, h->use_tma, h->prefetch_state, h->egress_strip, h->ingest_warps, h->band_from_state, h->profile ? 1 : h->lane_groups, h->analysis_only};
    mc_status st = MC_OK;
    switch (p->mode) {
        case MC_MODE_LAPLACE: st = h->motion.process(ctx, io, *p, levels, produced); break;
        case MC_MODE_COLOR: st = h->color.process(ctx, io, *p, levels, produced); break;
        case MC_MODE_PHASE: st = h->riesz.process(ctx, io, *p, levels, produced); break;
        default: break;
    }
    if (st != MC_OK) {
        // mirror the reference's recovery contract (ProcessingChain.cpp:50-62): state may be
        // half-updated, so drop it; the caller shows the input frame.
        reset_modes(h);
        tracker_reset(h);
        *produced = 0;
    }
    return st;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int mc_abi_version(void) { return MC_ABI_VERSION; }

int mc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void mc_params_default(mc_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->mode = MC_MODE_LAPLACE;
    p->levels = 4;
    p->framerate = 30.0;
    p->pre_downscale = 1;
    p->pre_roiW = 1.0f;
    p->pre_roiH = 1.0f;
}

static double hz_to_blend(double hz, double fps) {  // MagnificationParamsUi.hpp:29-34
    if (fps <= 0.0) fps = 30.0;
    if (hz <= 0.0) return 0.0;
    const double a = 1.0 - std::exp(-6.283185307179586 * hz / fps);
    return std::min(std::max(a, 0.0), 0.999999);
}

void mc_params_from_ui(mc_params* p, int mode, int amplification, double wavelength, double low_hz, double high_hz,
                       int chroma, int levels, double fps) {
    if (!p) return;
    mc_params_default(p);
    p->mode = mode;
    p->amplification = amplification;
    p->levels = levels;
    p->framerate = fps;
    switch (mode) {
        case MC_MODE_COLOR:
            p->coLow = low_hz; p->coHigh = high_hz;
            break;
        case MC_MODE_LAPLACE:
            p->coWavelength = wavelength * 10.0;
            p->coLow = hz_to_blend(low_hz, fps);
            p->coHigh = hz_to_blend(high_hz, fps);
            p->chromAttenuation = chroma / 100.0;
            break;
        case MC_MODE_PHASE:
            p->coWavelength = 100.0 - wavelength;
            p->coLow = low_hz; p->coHigh = high_hz;
            break;
        default: break;
    }
}

int mc_calculate_max_levels(int width, int height) { return calculate_max_levels(width, height); }
int mc_optimal_buffer_size(int fps) { return optimal_buffer_size(fps); }

mc_status mc_butterworth(unsigned order, double wn, double* a, double* b) try {
    debug_maybe_throw();
    if (!a || !b || order == 0 || order > 16) return MC_ERR_INVALID;
    std::vector<double> va, vb;
    butterworth(order, wn, va, vb);
    for (unsigned i = 0; i <= order; ++i) { a[i] = va[i]; b[i] = vb[i]; }
    return MC_OK;
} catch (...) { return on_exception(nullptr); }

mc_status mc_motion_gains(const mc_params* p, int levels, int width, int height, float* gains) try {
    debug_maybe_throw();
    if (!p || !gains || levels < 1) return MC_ERR_INVALID;
    std::vector<float> g;
    motion_gains(p->amplification, p->coWavelength, levels, width, height, g);
    for (int i = 0; i <= levels; ++i) gains[i] = g[(size_t)i];
    return MC_OK;
} catch (...) { return on_exception(nullptr); }

mc_status mc_create_lanes(int device, int lanes, mc_handle** out) try {
    // grid.z carries lanes * channels (<= 65535) and the Color min/max scratch is sized per lane
    if (!out || lanes < 1 || lanes > MC_MAX_LANES) { g_create_error = "bad arguments (1 <= lanes <= MC_MAX_LANES)"; return MC_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) {
        cudaGetLastError();
        g_create_error = "no usable CUDA device (this core has no CPU fallback)";
        return MC_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        cudaGetLastError();
        g_create_error = "device is not sm_100-class; kernels are built for sm_100a only";
        return MC_ERR_NO_DEVICE;
    }
    debug_maybe_throw();
    mc_handle* h = new mc_handle();
    try {
    h->device = device;
    h->lanes = lanes;
    auto fail = [&](const char* what, cudaError_t e) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        mc_destroy(h);
        return MC_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
    std::vector<LabLutCell> lut;
    build_lab_lut_cells(lut);
    if (lut.empty()) { g_create_error = "embedded Lab LUT missing"; mc_destroy(h); return MC_ERR_INVALID; }
    std::vector<float4> gam;
    build_inv_gamma_spline(gam);
    build_lab_inv_coeffs(h->tables.inv_coeffs);
    if ((e = cudaMalloc((void**)&h->tables.lab_lut, lut.size() * sizeof(LabLutCell))) != cudaSuccess) return fail("cudaMalloc lut", e);
    if ((e = cudaMalloc((void**)&h->tables.inv_gamma, gam.size() * sizeof(float4))) != cudaSuccess) return fail("cudaMalloc gamma", e);
    if ((e = cudaMemcpy(h->tables.lab_lut, lut.data(), lut.size() * sizeof(LabLutCell), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy lut", e);
    if ((e = cudaMemcpy(h->tables.inv_gamma, gam.data(), gam.size() * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy gamma", e);
    h->motion.lanes = h->color.lanes = h->riesz.lanes = lanes;
    *out = h;
    return MC_OK;
    } catch (...) { mc_destroy(h); throw; }
} catch (...) { return on_exception(nullptr); }

mc_status mc_create(int device, mc_handle** out) { return mc_create_lanes(device, 1, out); }

void mc_destroy(mc_handle* h) try {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->s_in) cudaStreamSynchronize(h->s_in);
    if (h->s_out) cudaStreamSynchronize(h->s_out);
    reset_modes(h);
    free_slots(h);
    if (h->float_out) cudaFree(h->float_out);
    if (h->c_raw) cudaFree(h->c_raw);
    if (h->c_pre) cudaFree(h->c_pre);
    if (h->c_gray) cudaFree(h->c_gray);
    if (h->c_out) cudaFree(h->c_out);
    if (h->c_tabs) cudaFree(h->c_tabs);
    if (h->tables.lab_lut) cudaFree(h->tables.lab_lut);
    if (h->tables.inv_gamma) cudaFree(h->tables.inv_gamma);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
} catch (...) {}

/* test hook, not declared in the public header: the n-th guarded entry on this thread throws std::bad_alloc */
void mc_debug_inject_exception(int nth) { g_inject_throw = nth; }

mc_status mc_reset(mc_handle* h) try {
    if (!h) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    reset_modes(h);
    tracker_reset(h);
    return MC_OK;
} catch (...) { return on_exception(h); }

mc_status mc_set_option(mc_handle* h, const char* key, int value) try {
    if (!h || !key) return MC_ERR_INVALID;
    if (!std::strcmp(key, "faithful_level0")) { h->faithful0 = value != 0; return MC_OK; }
    if (!std::strcmp(key, "keep_float_output")) { h->keep_float = value != 0; return MC_OK; }
    if (!std::strcmp(key, "profile_kernels")) { h->profile = value != 0; return MC_OK; }
    if (!std::strcmp(key, "use_tma")) { h->use_tma = value != 0; return MC_OK; }
    if (!std::strcmp(key, "prefetch_state")) { h->prefetch_state = value != 0; return MC_OK; }
    if (!std::strcmp(key, "lane_groups")) { h->lane_groups = value < 0 ? 0 : value; return MC_OK; }
    if (!std::strcmp(key, "egress_strip")) { h->egress_strip = value == 1 ? 20 : value; return MC_OK; }
    if (!std::strcmp(key, "ingest_warps")) { h->ingest_warps = value; return MC_OK; }
    if (!std::strcmp(key, "band_from_state")) { h->band_from_state = value != 0; return MC_OK; }
    if (!std::strcmp(key, "analysis_only")) { h->analysis_only = value != 0; return MC_OK; }
    if (!std::strcmp(key, "pipeline_depth")) {
        if (value < 1 || value > 16 || !h->inflight.empty()) { h->err = "bad pipeline_depth"; return MC_ERR_INVALID; }
        h->depth = value;
        free_slots(h);
        return MC_OK;
    }
    h->err = std::string("unknown option ") + key;
    return MC_ERR_INVALID;
} catch (...) { return on_exception(h); }

void* mc_stream(mc_handle* h) { return h ? (void*)h->stream : nullptr; }
uint64_t mc_launch_count(mc_handle* h) { return h ? h->launches : 0; }
int mc_pipeline_depth(mc_handle* h) { return h ? h->depth : 0; }

const char* mc_last_error(mc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

mc_status mc_sync(mc_handle* h) try {
    if (!h) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return MC_OK;
} catch (...) { return on_exception(h); }

mc_status mc_process_device(mc_handle* h, const uint8_t* d_in, int width, int height, int channels, size_t in_step,
                            const mc_params* p, uint8_t* d_out, size_t out_step, int* produced) try {
    if (!h || !produced) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    return process_device_impl(h, d_in, width, height, channels, in_step, p, d_out, out_step, produced);
} catch (...) { return on_exception(h); }

// submit with the destination known up front: when `in`/`out` are pinned (cudaHostAlloc /
// cudaHostRegister) the copies go straight between the caller's buffers and HBM (no staging memcpy).
mc_status mc_submit(mc_handle* h, const uint8_t* in, int width, int height, int channels, size_t in_step,
                    const mc_params* p, uint8_t* out, size_t out_step) try {
    if (!h || !p) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if ((int)h->inflight.size() >= h->depth) { h->err = "pipeline full: call mc_collect first"; return MC_ERR_INVALID; }
    const bool have = in != nullptr && width > 0 && height > 0 && (channels == 1 || channels == 3);
    const size_t row = have ? (size_t)width * channels : 0;
    const size_t bytes = row * (size_t)(have ? height : 0) * h->lanes;
    if (have && in_step < row) { h->err = "step too small"; return MC_ERR_INVALID; }
    mc_status st = ensure_slots(h, std::max<size_t>(bytes, 1));
    if (st != MC_OK) return st;
    const int si = h->next_slot;
    Slot& s = h->slots[(size_t)si];
    s.w = width; s.h = height; s.c = channels; s.produced = 0;
    s.user_out = out; s.user_out_step = out_step; s.direct_out = false;
    if (have && out && out_step < row) { h->err = "out_step too small"; return MC_ERR_INVALID; }
    if (!have) {
        int produced = 0;
        st = process_device_impl(h, nullptr, 0, 0, channels, 0, p, nullptr, 0, &produced);
        if (st != MC_OK) return st;
        CK(cudaEventRecord(s.ev_done, h->stream));
        h->inflight.push_back(si);
        h->next_slot = (si + 1) % h->depth;
        return MC_OK;
    }
    const size_t rows = (size_t)height * h->lanes;
    if (is_pinned(in)) {
        if (in_step == row) CK(cudaMemcpyAsync(s.d_in, in, bytes, cudaMemcpyHostToDevice, h->s_in));
        else CK(cudaMemcpy2DAsync(s.d_in, row, in, in_step, row, rows, cudaMemcpyHostToDevice, h->s_in));
    } else {
        for (size_t r = 0; r < rows; ++r) std::memcpy(s.h_in + r * row, in + r * in_step, row);
        CK(cudaMemcpyAsync(s.d_in, s.h_in, bytes, cudaMemcpyHostToDevice, h->s_in));
    }
    CK(cudaEventRecord(s.ev_in, h->s_in));
    CK(cudaStreamWaitEvent(h->stream, s.ev_in, 0));
    int produced = 0;
    st = process_device_impl(h, s.d_in, width, height, channels, row, p, s.d_out, row, &produced);
    if (st != MC_OK) return st;
    s.produced = produced;
    CK(cudaEventRecord(s.ev_k, h->stream));
    if (produced) {
        CK(cudaStreamWaitEvent(h->s_out, s.ev_k, 0));
        if (out && is_pinned(out)) {
            if (out_step == row) CK(cudaMemcpyAsync(out, s.d_out, bytes, cudaMemcpyDeviceToHost, h->s_out));
            else CK(cudaMemcpy2DAsync(out, out_step, s.d_out, row, row, rows, cudaMemcpyDeviceToHost, h->s_out));
            s.direct_out = true;
        } else if (out) {
            CK(cudaMemcpyAsync(s.h_out, s.d_out, bytes, cudaMemcpyDeviceToHost, h->s_out));
        }
        CK(cudaEventRecord(s.ev_done, h->s_out));
    } else {
        CK(cudaEventRecord(s.ev_done, h->stream));
    }
    // d_in of this slot must not be overwritten before its kernels ran: the next H2D into this slot
    // happens only after collect() of this frame, which waits on ev_done (>= ev_k).
    h->inflight.push_back(si);
    h->next_slot = (si + 1) % h->depth;
    return MC_OK;
} catch (...) { return on_exception(h); }

mc_status mc_collect(mc_handle* h, int* produced) try {
    if (!h || !produced) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (h->inflight.empty()) { h->err = "nothing in flight"; return MC_ERR_INVALID; }
    const int si = h->inflight.front();
    h->inflight.pop_front();
    Slot& s = h->slots[(size_t)si];
    CK(cudaEventSynchronize(s.ev_done));
    *produced = s.produced;
    if (s.produced && !s.direct_out && s.user_out) {
        const size_t row = (size_t)s.w * s.c, rows = (size_t)s.h * h->lanes;
        for (size_t r = 0; r < rows; ++r) std::memcpy(s.user_out + r * s.user_out_step, s.h_out + r * row, row);
    }
    return MC_OK;
} catch (...) { return on_exception(h); }

mc_status mc_process(mc_handle* h, const uint8_t* in, int width, int height, int channels, size_t in_step,
                     const mc_params* p, uint8_t* out, size_t out_step, int* produced) try {
    if (!h || !produced) return MC_ERR_INVALID;
    *produced = 0;
    if (!h->inflight.empty()) { h->err = "mc_process called with pipelined frames in flight"; return MC_ERR_INVALID; }
    mc_status st = mc_submit(h, in, width, height, channels, in_step, p, out, out_step);
    if (st != MC_OK) return st;
    return mc_collect(h, produced);
} catch (...) { return on_exception(h); }

mc_status mc_state_dims(mc_handle* h, const char* name, int level, int* rows, int* cols, int* channels) try {
    if (!h || !name || !rows || !cols || !channels) return MC_ERR_INVALID;
    *rows = *cols = *channels = 0;
    StateRef r;
    if (h->t_mode == MC_MODE_LAPLACE) h->motion.find_state(name, level, r);
    else if (h->t_mode == MC_MODE_COLOR) h->color.find_state(name, level, r);
    else if (h->t_mode == MC_MODE_PHASE) h->riesz.find_state(name, level, r);
    if (r.ptr) { *rows = r.rows; *cols = r.cols; *channels = r.channels; }
    return MC_OK;
} catch (...) { return on_exception(h); }

static mc_status state_xfer(mc_handle* h, const char* name, int level, float* host, size_t n, bool get) try {
    if (!h || !name || !host) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    StateRef r;
    if (h->t_mode == MC_MODE_LAPLACE) h->motion.find_state(name, level, r);
    else if (h->t_mode == MC_MODE_COLOR) h->color.find_state(name, level, r);
    else if (h->t_mode == MC_MODE_PHASE) h->riesz.find_state(name, level, r);
    if (!r.ptr) { h->err = std::string("no such state: ") + name; return MC_ERR_INVALID; }
    const size_t planes = (size_t)h->lanes * r.channels;
    if (n < planes * r.rows * r.cols) { h->err = "state buffer too small"; return MC_ERR_INVALID; }
    CK(cudaStreamSynchronize(h->stream));
    for (size_t pl = 0; pl < planes; ++pl) {
        float* d = r.ptr + pl * r.plane_stride;
        float* hp = host + pl * (size_t)r.rows * r.cols;
        if (get) CK(cudaMemcpy2D(hp, (size_t)r.cols * 4, d, (size_t)r.pitch * 4, (size_t)r.cols * 4, r.rows, cudaMemcpyDeviceToHost));
        else CK(cudaMemcpy2D(d, (size_t)r.pitch * 4, hp, (size_t)r.cols * 4, (size_t)r.cols * 4, r.rows, cudaMemcpyHostToDevice));
    }
    return MC_OK;
} catch (...) { return on_exception(h); }

mc_status mc_get_state(mc_handle* h, const char* name, int level, float* dst, size_t n) try {
    return state_xfer(h, name, level, dst, n, true);
} catch (...) { return on_exception(h); }
mc_status mc_set_state(mc_handle* h, const char* name, int level, const float* src, size_t n) try {
    return state_xfer(h, name, level, const_cast<float*>(src), n, false);
} catch (...) { return on_exception(h); }

mc_status mc_get_float_output(mc_handle* h, float* dst, size_t n) try {
    if (!h || !dst) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->float_out || n < h->float_out_floats) { h->err = "no float output kept (set keep_float_output) or buffer too small"; return MC_ERR_INVALID; }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(dst, h->float_out, h->float_out_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return MC_OK;
} catch (...) { return on_exception(h); }

}  // extern "C"

extern "C" void* mc_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void mc_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" mc_status mc_profile_read(mc_handle* h, char* buf, size_t cap) try {
    if (!h || !buf || cap == 0) return MC_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    struct Acc { std::string name; int level; int n; double ms; };
    std::vector<Acc> acc;
    for (auto& r : h->prof.recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
        bool found = false;
        for (auto& a : acc)
            if (a.level == r.level && a.name == r.name) { a.n++; a.ms += ms; found = true; break; }
        if (!found) acc.push_back(Acc{r.name, r.level, 1, ms});
    }
    h->prof.recs.clear();
    std::string out;
    char line[160];
    for (auto& a : acc) {
        std::snprintf(line, sizeof(line), "%s %d %d %.6f\n", a.name.c_str(), a.level, a.n, a.ms);
        out += line;
    }
    if (out.size() + 1 > cap) { h->err = "profile buffer too small"; return MC_ERR_INVALID; }
    std::memcpy(buf, out.c_str(), out.size() + 1);
    return MC_OK;
} catch (...) { return on_exception(h); }

namespace {
mc_status grow(mc_handle* h, uint8_t** p, size_t* cap, size_t need) {
    if (need <= *cap) return MC_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CK(cudaMalloc((void**)p, need));
    *cap = need;
    return MC_OK;
}
}  // namespace

// runChainOnce (ChainBuilder.cpp:19-29) = Preprocess -> Grayscale -> Magnification, fused on the device.
extern "C" mc_status mc_chain_process(mc_handle* h, const uint8_t* in, int width, int height, int channels, size_t in_step,
                                      const mc_params* p, int grayscale, uint8_t* out, size_t out_bytes, uint8_t* original,
                                      size_t original_bytes, mc_chain_info* info) try {
    if (!h || !p || !info) return MC_ERR_INVALID;
    std::memset(info, 0, sizeof(*info));
    info->cur_is_input = 1; info->orig_is_input = 1;
    CK(cudaSetDevice(h->device));
    if (h->lanes != 1) { h->err = "mc_chain_process needs a 1-lane handle"; return MC_ERR_INVALID; }
    if (!h->inflight.empty()) { h->err = "mc_chain_process called with pipelined frames in flight"; return MC_ERR_INVALID; }
    int produced = 0;
    if (in == nullptr || width <= 0 || height <= 0) {   // empty image: every stage is an identity (PreprocessProcessor.cpp:11)
        return process_device_impl(h, nullptr, 0, 0, channels, 0, p, nullptr, 0, &produced);
    }
    if (channels != 1 && channels != 3) { h->err = "channels must be 1 or 3"; return MC_ERR_INVALID; }
    const size_t row = (size_t)width * channels;
    if (in_step < row) { h->err = "step too small"; return MC_ERR_INVALID; }
    mc_status st;
    if ((st = grow(h, &h->c_raw, &h->c_raw_b, row * height)) != MC_OK) return st;
    CK(cudaMemcpy2DAsync(h->c_raw, row, in, in_step, row, (size_t)height, cudaMemcpyHostToDevice, h->stream));

    // ---- PreprocessProcessor::process (PreprocessProcessor.cpp:10-51)
    const int divisor = std::min(std::max((int)p->pre_downscale, 1), 8);
    const bool pre_active = p->pre_roiEnabled != 0 || divisor != 1;
    const uint8_t* cur = h->c_raw;
    int cw = width, chh = height, cc = channels;
    if (pre_active) {
        int rx, ry, rw, rh;
        preprocess_roi(width, height, p->pre_roiEnabled != 0, p->pre_roiX, p->pre_roiY, p->pre_roiW, p->pre_roiH, rx, ry, rw, rh);
        const int dw = divisor > 1 ? std::max(1, rw / divisor) : rw, dh = divisor > 1 ? std::max(1, rh / divisor) : rh;
        if ((st = grow(h, &h->c_pre, &h->c_pre_b, (size_t)dw * dh * channels)) != MC_OK) return st;
        const AreaTap *dxt = nullptr, *dyt = nullptr;
        const int *dxo = nullptr, *dyo = nullptr;
        if (divisor > 1) {
            std::vector<AreaTap> xt, yt;
            std::vector<int> xo, yo;
            build_area_tab(rw, dw, (double)rw / dw, xt, xo);
            build_area_tab(rh, dh, (double)rh / dh, yt, yo);
            const size_t bx = xt.size() * sizeof(AreaTap), by = yt.size() * sizeof(AreaTap), box = xo.size() * sizeof(int), boy = yo.size() * sizeof(int);
            uint8_t* tb = (uint8_t*)h->c_tabs;
            size_t tcap = h->c_tabs_b;
            if ((st = grow(h, &tb, &tcap, bx + by + box + boy + 64)) != MC_OK) return st;
            h->c_tabs = tb; h->c_tabs_b = tcap;
            // the tables are small (a few KB); synchronous copies keep their host vectors alive long enough
            CK(cudaStreamSynchronize(h->stream));
            CK(cudaMemcpy(tb, xt.data(), bx, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(tb + bx, yt.data(), by, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(tb + bx + by, xo.data(), box, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(tb + bx + by + box, yo.data(), boy, cudaMemcpyHostToDevice));
            dxt = (const AreaTap*)tb; dyt = (const AreaTap*)(tb + bx);
            dxo = (const int*)(tb + bx + by); dyo = (const int*)(tb + bx + by + box);
        }
        uint8_t* gray = nullptr;
        if (grayscale && channels == 3) {
            if ((st = grow(h, &h->c_gray, &h->c_gray_b, (size_t)dw * dh)) != MC_OK) return st;
            gray = h->c_gray;
        }
        CK(launch_preprocess(h->c_raw + (size_t)ry * row + (size_t)rx * channels, row, channels, rw, rh, dw, dh, divisor == 1,
                             dxt, dxo, dyt, dyo, h->c_pre, gray, h->stream));
        ++h->launches;
        cur = h->c_pre; cw = dw; chh = dh;
        info->orig_is_input = 0; info->orig_w = dw; info->orig_h = dh; info->orig_channels = channels;
        if (original) {
            if (original_bytes < (size_t)dw * dh * channels) { h->err = "original buffer too small"; return MC_ERR_INVALID; }
            CK(cudaMemcpyAsync(original, h->c_pre, (size_t)dw * dh * channels, cudaMemcpyDeviceToHost, h->stream));
        }
        if (gray) { cur = gray; cc = 1; }
    } else if (grayscale && channels == 3) {
        // ---- GrayscaleProcessor::process alone (GrayscaleProcessor.cpp:7-16)
        if ((st = grow(h, &h->c_gray, &h->c_gray_b, (size_t)width * height)) != MC_OK) return st;
        CK(launch_preprocess(h->c_raw, row, 3, width, height, width, height, true, nullptr, nullptr, nullptr, nullptr, nullptr,
                             h->c_gray, h->stream));
        ++h->launches;
        cur = h->c_gray; cc = 1;
    }
    const bool cur_is_raw = cur == h->c_raw;

    // ---- MagnificationProcessor::process on the chain's current frame
    const size_t crow = (size_t)cw * cc;
    if ((st = grow(h, &h->c_out, &h->c_out_b, crow * chh)) != MC_OK) return st;
    st = process_device_impl(h, cur, cw, chh, cc, crow, p, h->c_out, crow, &produced);
    if (st != MC_OK) return st;
    info->magnified = produced;
    const uint8_t* result = produced ? h->c_out : cur;
    if (produced || !cur_is_raw) {
        info->cur_is_input = 0; info->out_w = cw; info->out_h = chh; info->out_channels = cc;
        if (out) {
            if (out_bytes < crow * chh) { h->err = "out buffer too small"; return MC_ERR_INVALID; }
            CK(cudaMemcpyAsync(out, result, crow * chh, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    return MC_OK;
} catch (...) { return on_exception(h); }
