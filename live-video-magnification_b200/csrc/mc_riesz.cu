// Phase (Riesz) mode — placeholder until the Riesz pyramid path lands.
#include "mc_modes.h"
namespace mc {
void RieszMode::reset() { arena.release(); allocated = false; }
mc_status RieszMode::process(const ModeCtx& ctx, const FrameIO&, const mc_params&, int, int* produced) {
    *produced = 0;
    *ctx.err = "Phase mode not implemented yet";
    return MC_ERR_UNSUPPORTED;
}
void RieszMode::find_state(const char*, int, StateRef& out) { out = StateRef{}; }
}  // namespace mc
