// TEST INFRASTRUCTURE ONLY (oracle/).  Lets the drop-in check inside _livim_ref use the reference-side adapter
// (live-video-magnification_b200/adapter/MagnificationProcessorB200.hpp) without making the checker module depend
// on the product library at load time: the six C-ABI entry points the adapter calls are resolved with dlopen/dlsym
// from the path handed to mc_dl_open() (libmagcore_b200.so), on first use.
#include <dlfcn.h>

#include <stdexcept>
#include <string>

#include "magcore_b200.h"

namespace {
void* g_lib = nullptr;
std::string g_path;

template <typename F> F sym(const char* name) {
    if (!g_lib) {
        if (g_path.empty()) throw std::runtime_error("mc_dl: library path not set (call set_magcore_library first)");
        g_lib = dlopen(g_path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!g_lib) throw std::runtime_error(std::string("mc_dl: ") + dlerror());
    }
    void* p = dlsym(g_lib, name);
    if (!p) throw std::runtime_error(std::string("mc_dl: missing symbol ") + name);
    return reinterpret_cast<F>(p);
}
}  // namespace

void mc_dl_open(const std::string& path) { g_path = path; }

extern "C" {
__attribute__((visibility("hidden"))) void mc_params_default(mc_params* p) { sym<void (*)(mc_params*)>("mc_params_default")(p); }
__attribute__((visibility("hidden"))) mc_status mc_create(int device, mc_handle** out) {
    return sym<mc_status (*)(int, mc_handle**)>("mc_create")(device, out);
}
__attribute__((visibility("hidden"))) void mc_destroy(mc_handle* h) {
    if (h) sym<void (*)(mc_handle*)>("mc_destroy")(h);
}
__attribute__((visibility("hidden"))) mc_status mc_reset(mc_handle* h) { return sym<mc_status (*)(mc_handle*)>("mc_reset")(h); }
__attribute__((visibility("hidden"))) mc_status mc_process(mc_handle* h, const uint8_t* in, int width, int height, int channels,
                                                           size_t in_step, const mc_params* p, uint8_t* out, size_t out_step,
                                                           int* produced) {
    return sym<mc_status (*)(mc_handle*, const uint8_t*, int, int, int, size_t, const mc_params*, uint8_t*, size_t, int*)>(
        "mc_process")(h, in, width, height, channels, in_step, p, out, out_step, produced);
}
__attribute__((visibility("hidden"))) const char* mc_last_error(mc_handle* h) {
    return sym<const char* (*)(mc_handle*)>("mc_last_error")(h);
}
}
