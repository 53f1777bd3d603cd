// Motion (Laplace) per-frame driver — device twin of magcore::magnifyMotion
// (reference src/processing/magnification/MagnifyCore.hpp:83-160).
#include <cstring>

#include "mc_modes.h"

namespace mc {

cudaError_t DeviceArena::alloc(float** p, size_t floats) {
    void* v = nullptr;
    cudaError_t e = alloc_bytes(&v, floats * sizeof(float));
    *p = (float*)v;
    return e;
}
cudaError_t DeviceArena::alloc_bytes(void** p, size_t bytes) {
    *p = nullptr;
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 4);
    if (e == cudaSuccess) blocks.push_back(*p);
    return e;
}
void DeviceArena::release() {
    for (void* b : blocks) cudaFree(b);
    blocks.clear();
}

void MotionMode::reset() {
    arena.release();
    lv.clear(); G.clear(); hi.clear(); lo.clear(); M.clear();
    allocated = false;
    empty = true;
}

mc_status MotionMode::allocate(const ModeCtx& ctx, const FrameIO& io, int nlevels) {
    reset();
    levels = nlevels; channels = io.channels; w = io.w; h = io.h;
    faithful = ctx.faithful0;
    const size_t planes = (size_t)lanes * channels;
    lv.resize((size_t)levels + 1);
    int cw = w, ch = h;
    for (int l = 0; l <= levels; ++l) {
        lv[(size_t)l] = make_level(cw, ch);
        cw = (cw + 1) / 2;
        ch = (ch + 1) / 2;
    }
    G.assign((size_t)levels + 1, nullptr);
    hi.assign((size_t)levels + 1, nullptr);
    lo.assign((size_t)levels + 1, nullptr);
    M.assign((size_t)levels + 1, nullptr);
    for (int l = 0; l <= levels; ++l) {
        const size_t n = planes * lv[(size_t)l].plane;
        const bool band_level = l < levels;                       // bands 0..levels-1, residual = levels
        const bool live = band_level && l >= 1;                   // bands whose gain can be non-zero
        if (l >= 1 || faithful) MCK(arena.alloc(&G[(size_t)l], n));
        if (live || faithful) {
            MCK(arena.alloc(&hi[(size_t)l], n));
            MCK(arena.alloc(&lo[(size_t)l], n));
        }
        if (live) MCK(arena.alloc(&M[(size_t)l], n));
    }
    allocated = true;
    return MC_OK;
}

mc_status MotionMode::process(const ModeCtx& ctx, const FrameIO& io, const mc_params& p, int nlevels, int* produced) {
    if (!allocated || faithful != ctx.faithful0) {
        mc_status st = allocate(ctx, io, nlevels);
        if (st != MC_OK) return st;
    }
    const int planes = lanes * channels;
    const bool first = empty;  // MagnifyCore.hpp:98

    std::vector<float> gains;
    motion_gains(p.amplification, p.coWavelength, levels, w, h, gains);
    double c_lo = p.coLow, c_hi = p.coHigh;
    if (c_lo == 0) c_lo = 0.01;  // TemporalFilter.cpp:11-12

    // analysis: level 0 (u8 -> Lab/gray -> pyrDown)
    const int l_start = faithful ? 0 : 1;
    if (levels >= 2 || faithful)
        LAUNCH("ingest_down", 0, launch_ingest_down(io, *ctx.tables, lv[0], lv[1], G[1], faithful ? G[0] : nullptr, ctx.stream));
    // analysis: one fused kernel per level
    for (int l = l_start; l < levels; ++l) {
        LevelArgs a;
        a.lf = lv[(size_t)l]; a.lc = lv[(size_t)l + 1];
        a.g = G[(size_t)l]; a.g_next = G[(size_t)l + 1];
        a.hi = hi[(size_t)l]; a.lo = lo[(size_t)l];
        a.m = first ? nullptr : M[(size_t)l];
        a.planes = planes;
        a.first = first ? 1 : 0;
        a.c_hi = c_hi; a.one_minus_c_hi = 1 - c_hi; a.c_lo = c_lo; a.one_minus_c_lo = 1 - c_lo;
        a.gain = gains[(size_t)l];
        LAUNCH("level", l, launch_level(a, ctx.stream));
    }
    if (first && faithful)  // st.lowpassHi/Lo[levels] = residual (MagnifyCore.hpp:100-101)
    {
        const size_t n = (size_t)planes * lv[(size_t)levels].plane;
        LAUNCH("copy", levels, launch_copy_planes(hi[(size_t)levels], G[(size_t)levels], n, ctx.stream));
        LAUNCH("copy", levels, launch_copy_planes(lo[(size_t)levels], G[(size_t)levels], n, ctx.stream));
    }
    const float* m1 = nullptr;
    if (!first && levels >= 2) {
        // synthesis: residual and finest band are zero (MagnifyCore.hpp:130-131), so the collapse
        // starts from band levels-1 and stops at level 1; level 0 is folded into egress.
        for (int l = levels - 2; l >= 1; --l)
            LAUNCH("collapse", l, launch_collapse(lv[(size_t)l], lv[(size_t)l + 1], M[(size_t)l], M[(size_t)l + 1], planes, ctx.stream));
        m1 = M[1];
    }
    LAUNCH("egress", 0, launch_egress(io, *ctx.tables, lv[0], lv[levels >= 1 ? 1 : 0], m1, (float)p.chromAttenuation, ctx.float_out, ctx.stream));
    empty = false;
    *produced = 1;
    return MC_OK;
}

void MotionMode::find_state(const char* name, int level, StateRef& out) {
    out = StateRef{};
    if (!allocated || empty || level < 0 || level > levels) return;
    float* p = nullptr;
    if (!std::strcmp(name, "lowpassHi")) p = hi[(size_t)level];
    else if (!std::strcmp(name, "lowpassLo")) p = lo[(size_t)level];
    if (!p) return;
    const Level& l = lv[(size_t)level];
    out.ptr = p; out.rows = l.h; out.cols = l.w; out.channels = channels; out.pitch = l.pitch; out.plane_stride = l.plane;
}

}  // namespace mc
