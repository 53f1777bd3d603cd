// Motion (Laplace) per-frame driver — device twin of magcore::magnifyMotion
// (reference src/processing/magnification/MagnifyCore.hpp:83-160).
#include <algorithm>
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
void DeviceArena::free_block(void* p) {
    for (size_t i = 0; i < blocks.size(); ++i)
        if (blocks[i] == p) {
            cudaFree(p);
            blocks.erase(blocks.begin() + (long)i);
            return;
        }
}
void DeviceArena::release() {
    for (void* b : blocks) cudaFree(b);
    blocks.clear();
}

void MotionMode::reset() {
    arena.release();
    lv.clear(); G.clear(); hi.clear(); lo.clear(); M.clear();
    lab16 = nullptr;
    allocated = false;
    empty = true;
}

mc_status MotionMode::allocate(const ModeCtx& ctx, const FrameIO& io, int nlevels) {
    reset();
    levels = nlevels; channels = io.channels; w = io.w; h = io.h;
    faithful = ctx.faithful0;
    from_state = ctx.band_from_state;
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
        if (l >= 1) MCK(arena.alloc(&G[(size_t)l], n));
        if (live || faithful) {
            MCK(arena.alloc(&hi[(size_t)l], n));
            MCK(arena.alloc(&lo[(size_t)l], n));
        }
        // M_l: the amplified band of level l, overwritten by the collapsed cur_l.  With band_from_state the bands are
        // rebuilt from hi/lo by their consumers and only cur_2 .. cur_{levels-2} are materialised.
        if (live && (!from_state || (l >= 2 && l <= levels - 2))) MCK(arena.alloc(&M[(size_t)l], n));
    }
    // TMA descriptors for the f32 inputs of the fused level kernels
    tmaps.assign((size_t)levels + 1, TensorMapStorage{});
    tmap_valid.assign((size_t)levels + 1, 0);
    tmaps_hi.assign((size_t)levels + 1, TensorMapStorage{});
    tmaps_lo.assign((size_t)levels + 1, TensorMapStorage{});
    for (int l = 1; l < levels; ++l) {
        const Level& L = lv[(size_t)l];
        tmap_valid[(size_t)l] = make_level_tensor_map(&tmaps[(size_t)l], G[(size_t)l], L, (int)planes) &&
                                make_level_tensor_map(&tmaps_hi[(size_t)l], hi[(size_t)l], L, (int)planes, true) &&
                                make_level_tensor_map(&tmaps_lo[(size_t)l], lo[(size_t)l], L, (int)planes, true) ? 1 : 0;
    }
    if (channels == 3) {
        pitch16 = round_up(w, 64);
        plane16 = (size_t)h * pitch16;
        void* p = nullptr;
        MCK(arena.alloc_bytes(&p, planes * plane16 * sizeof(int16_t)));
        lab16 = (int16_t*)p;
    }
    allocated = true;
    return MC_OK;
}

mc_status MotionMode::process(const ModeCtx& ctx, const FrameIO& io, const mc_params& p, int nlevels, int* produced) {
    if (!allocated || faithful != ctx.faithful0 || from_state != ctx.band_from_state) {
        mc_status st = allocate(ctx, io, nlevels);
        if (st != MC_OK) return st;
    }
    const int planes = lanes * channels;
    const bool first = empty;  // MagnifyCore.hpp:98

    motion_gains(p.amplification, p.coWavelength, levels, w, h, gains);
    double c_lo = p.coLow, c_hi = p.coHigh;
    if (c_lo == 0) c_lo = 0.01;  // TemporalFilter.cpp:11-12

    // ingest: u8 BGR -> Lab16 planes (gray frames are read directly by the level-0 kernel)
    // (production path, >= 2 levels: one fused kernel also builds G1; otherwise Lab16 alone)
    const bool fused_ingest = channels == 3 && !faithful && levels >= 2;
    if (fused_ingest) LAUNCH("ingest_lab", 0, launch_ingest_lab(io, *ctx.tables, lab16, pitch16, plane16, G[1], lv[1], ctx.stream, ctx.ingest_warps));
    else if (channels == 3) LAUNCH("lab16", 0, launch_lab16(io, *ctx.tables, lab16, pitch16, plane16, ctx.stream));

    // analysis: one fused kernel per level (level 0 only builds G1 unless the faithful option is on)
    const int l_begin = fused_ingest ? 1 : ((levels >= 2 || faithful) ? 0 : levels);
    for (int l = l_begin; l < levels; ++l) {
        LevelArgs a;
        if (l == 0) {
            if (channels == 3) {
                a.in_kind = 1; a.g = lab16; a.in_plane = plane16; a.in_row = pitch16;
                a.sc[0] = 100.0f / 16384.0f; a.of[0] = 0.0f;
                a.sc[1] = a.sc[2] = 1.0f / 64.0f; a.of[1] = a.of[2] = -128.0f;
            } else {
                a.in_kind = 2; a.g = io.in; a.in_plane = io.in_lane_stride; a.in_row = (int)io.in_step;
                a.sc[0] = 0.003921568859368563f;
            }
        } else {
            a.in_kind = 0; a.g = G[(size_t)l]; a.in_plane = lv[(size_t)l].plane; a.in_row = lv[(size_t)l].pitch;
            if (tmap_valid[(size_t)l] && ctx.use_tma) {
                a.tmap = &tmaps[(size_t)l];
                if (ctx.prefetch_state) { a.tmap_hi = &tmaps_hi[(size_t)l]; a.tmap_lo = &tmaps_lo[(size_t)l]; }
            }
        }
        a.channels = channels;
        a.lf = lv[(size_t)l]; a.lc = lv[(size_t)l + 1];
        a.g_next = G[(size_t)l + 1];
        a.hi = hi[(size_t)l]; a.lo = lo[(size_t)l];
        a.m = (first || from_state) ? nullptr : M[(size_t)l];
        a.planes = planes;
        a.first = first ? 1 : 0;
        a.band = (l >= 1 || faithful) ? 1 : 0;
        a.c_hi = c_hi; a.one_minus_c_hi = 1 - c_hi; a.c_lo = c_lo; a.one_minus_c_lo = 1 - c_lo;
        a.gain = gains[(size_t)l];
        if (a.band) LAUNCH("level", l, launch_level(a, ctx.stream));
        else LAUNCH("down", l, launch_down(a, ctx.stream));
    }
    if (first && faithful)  // st.lowpassHi/Lo[levels] = residual (MagnifyCore.hpp:100-101)
    {
        const size_t n = (size_t)planes * lv[(size_t)levels].plane;
        LAUNCH("copy", levels, launch_copy_planes(hi[(size_t)levels], G[(size_t)levels], n, ctx.stream));
        LAUNCH("copy", levels, launch_copy_planes(lo[(size_t)levels], G[(size_t)levels], n, ctx.stream));
    }
    if (ctx.analysis_only && !first) {   // state-carry pass: the temporal state is up to date, no frame is produced
        *produced = 0;
        return MC_OK;
    }
    BandSrc m1, c2;
    if (!first && levels >= 2) {
        // synthesis: residual and finest band are zero (MagnifyCore.hpp:130-131), so the collapse starts from
        // band levels-1 (cur_{levels-1} = 0 + m_{levels-1}); levels 1 and 0 are folded into egress.
        auto band = [&](int l) {
            return from_state ? BandSrc{hi[(size_t)l], lo[(size_t)l], gains[(size_t)l]} : BandSrc{M[(size_t)l], nullptr, 1.0f};
        };
        auto cur = [&](int l) { return l == levels - 1 ? band(l) : BandSrc{M[(size_t)l], nullptr, 1.0f}; };
        for (int l = levels - 2; l >= 2; --l)
            LAUNCH("collapse", l, launch_collapse(lv[(size_t)l], lv[(size_t)l + 1], band(l), cur(l + 1), M[(size_t)l], planes, ctx.stream));
        m1 = band(1);
        if (levels >= 3) c2 = cur(2);
    }
    const Level& l1 = lv[levels >= 1 ? 1 : 0];
    const Level& l2 = lv[levels >= 2 ? 2 : 0];
    LAUNCH("egress", 0, launch_egress(io, *ctx.tables, lab16, pitch16, plane16, m1, l1, c2, l2, (float)p.chromAttenuation, ctx.float_out,
                                      ctx.stream, ctx.egress_strip));
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
