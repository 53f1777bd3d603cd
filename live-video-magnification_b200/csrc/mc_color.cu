// Color mode — placeholder until the Gaussian + ideal-FFT path lands.
#include "mc_modes.h"
namespace mc {
void ColorMode::reset() { arena.release(); allocated = false; count = 0; head = 0; }
mc_status ColorMode::process(const ModeCtx& ctx, const FrameIO&, const mc_params&, int, int* produced) {
    *produced = 0;
    *ctx.err = "Color mode not implemented yet";
    return MC_ERR_UNSUPPORTED;
}
void ColorMode::find_state(const char*, int, StateRef& out) { out = StateRef{}; }
}  // namespace mc
