/*
 * magcore_b200.h — C ABI of the B200-native Eulerian video-magnification core.
 *
 * Drop-in boundary: the reference's per-frame hot path sits behind
 *     livim::IProcessor::process(const FrameRef&, const ProcessorConfig&) / reset()
 *     (reference src/processing/IProcessor.hpp:50-60) as implemented by
 *     livim::MagnificationProcessor (src/processing/MagnificationProcessor.cpp:10-67).
 * One mc_handle == one MagnificationProcessor instance (it owns all temporal state, one CUDA
 * stream, no globals); `mc_process` has exactly that method's semantics on raw host pixels.
 *
 * Everything is extern "C", plain pointers and sizes; no exceptions cross this boundary (the C++
 * adapter in live-video-magnification_b200/adapter/ rethrows MC_ERR_* as std::runtime_error so the
 * reference's exception firewall, src/processing/ProcessingChain.cpp:50-62, keeps working).
 */
#ifndef MAGCORE_B200_H
#define MAGCORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_ABI_VERSION 2

typedef struct mc_handle mc_handle;

typedef enum mc_status {
    MC_OK = 0,
    MC_ERR_INVALID = 1,     /* bad argument */
    MC_ERR_CUDA = 2,        /* a CUDA runtime / cuFFT call failed; see mc_last_error() */
    MC_ERR_NO_DEVICE = 3,   /* no usable sm_100 device: the core has no CPU fallback */
    MC_ERR_UNSUPPORTED = 4,
    MC_ERR_INTERNAL = 5     /* a C++ exception (std::bad_alloc, ...) was caught at the boundary; the handle's temporal
                               state has been dropped, as after any failed frame; see mc_last_error() */
} mc_status;

#define MC_MAX_LANES 4096   /* upper bound of mc_create_lanes(): lanes * channels is a CUDA grid.z extent */

/* livim::MagnificationMode, src/processing/IProcessor.hpp:10 (same numeric order). */
typedef enum mc_mode { MC_MODE_LAPLACE = 0, MC_MODE_PHASE = 1, MC_MODE_COLOR = 2, MC_MODE_NONE = 3 } mc_mode;

/* POD copy of livim::MagnificationParams (IProcessor.hpp:14-23) plus the PreprocessParams
 * fingerprint (IProcessor.hpp:26-41) which only takes part in the structural-reset decision
 * (src/processing/magnification/MagnifyCore.hpp:53-65).  Units are ALGORITHM units: Laplace
 * coLow/coHigh are EMA blend coefficients, Color/Phase cutoffs are Hz. */
typedef struct mc_params {
    int32_t mode;              /* mc_mode */
    int32_t levels;
    double amplification;
    double coWavelength;
    double coLow;
    double coHigh;
    double chromAttenuation;
    double framerate;
    int32_t pre_downscale;     /* PreprocessParams::downscale */
    int32_t pre_roiEnabled;
    float pre_roiX, pre_roiY, pre_roiW, pre_roiH;
} mc_params;

/* Fills *p with the reference defaults (MagnificationParams{} + PreprocessParams{}). */
void mc_params_default(mc_params* p);

/* UI units -> algorithm units; replaces livim::toParams(), src/processing/MagnificationParamsUi.hpp:74-103
 * (amplification, wavelength %, low/high Hz, chroma %, levels, captureFps). */
void mc_params_from_ui(mc_params* p, int mode, int amplification, double wavelength, double low_hz,
                       double high_hz, int chroma, int levels, double fps);

/* livim::calculateMaxLevels, src/processing/magnification/SpatialFilter.cpp:5-11. */
int mc_calculate_max_levels(int width, int height);

/* livim::getOptimalBufferSize, src/processing/magnification/TemporalFilter.cpp:82-94. */
int mc_optimal_buffer_size(int fps);

/* livim::butterworth(N, Wn, a, b), src/processing/magnification/TemporalFilter.cpp:279-297;
 * a and b must hold N+1 doubles. */
mc_status mc_butterworth(unsigned order, double wn, double* a, double* b);

/* Per-level Laplace gains, src/processing/magnification/MagnifyCore.hpp:114-134; gains[levels+1]. */
mc_status mc_motion_gains(const mc_params* p, int levels, int width, int height, float* gains);

int mc_abi_version(void);
int mc_device_count(void);

/* Construct / destroy a processor bound to CUDA device `device` (replaces
 * std::make_unique<MagnificationProcessor>(), src/processing/ChainBuilder.cpp:15).
 * `lanes` >= 1 independent streams are stepped in lock-step by one handle (lanes == 1 is the
 * reference's processor; lanes > 1 is the throughput form: one launch set serves all lanes,
 * each lane keeping its own temporal state — used for multi-stream serving and for benchmarking
 * against a working set larger than L2). */
mc_status mc_create(int device, mc_handle** out);
mc_status mc_create_lanes(int device, int lanes, mc_handle** out);
void mc_destroy(mc_handle* h);

/* MagnificationProcessor::reset(), MagnificationProcessor.cpp:10-15. */
mc_status mc_reset(mc_handle* h);

/* MagnificationProcessor::process(), MagnificationProcessor.cpp:17-67, on host pixels.
 *   in   : `lanes` frames back to back, each h rows of `in_step` bytes, CV_8UC3 BGR interleaved
 *          (channels == 3) or CV_8UC1 (channels == 1); treated as immutable.
 *   out  : same geometry with `out_step`; written only when *produced != 0.
 *   *produced == 0  <=>  the reference returns the input FrameRef unchanged (mode None, empty or
 *          too-small image, Color warm-up, Riesz first frame / gray input).
 * Blocking: H2D copy, kernels and D2H copy complete before it returns. */
mc_status mc_process(mc_handle* h, const uint8_t* in, int width, int height, int channels,
                     size_t in_step, const mc_params* p, uint8_t* out, size_t out_step,
                     int* produced);

/* Same contract on DEVICE pointers (frames already resident in HBM; lane stride = h*step).
 * Work is enqueued on the handle's stream; call mc_sync() before reading d_out on another stream. */
mc_status mc_process_device(mc_handle* h, const uint8_t* d_in, int width, int height, int channels,
                            size_t in_step, const mc_params* p, uint8_t* d_out, size_t out_step,
                            int* produced);
mc_status mc_sync(mc_handle* h);

/* --- "next" row: the whole processing chain of one frame on the device ---------------------
 * Replaces runChainOnce(chain, in, cfg, original) (reference src/processing/ChainBuilder.cpp:19-29) over
 * PreprocessProcessor (ROI crop + INTER_AREA downscale, PreprocessProcessor.cpp:10-51), GrayscaleProcessor
 * (BGR2GRAY, GrayscaleProcessor.cpp:7-16) and MagnificationProcessor: the raw frame is uploaded once, both
 * front stages run bit-exact on the B200, the magnification core runs on their result, and the processed frame
 * plus the "original" tap (the pre-magnification frame, ChainBuilder.cpp:25) come back.  lanes must be 1.
 * `out` / `original` are written tight (step = width * channels); the *_is_input flags mirror the reference
 * returning the very same FrameRef (nothing written). */
typedef struct mc_chain_info {
    int32_t cur_is_input;                    /* processed frame == the input FrameRef */
    int32_t out_w, out_h, out_channels;      /* geometry of `out` when cur_is_input == 0 */
    int32_t orig_is_input;                   /* original tap == the input FrameRef */
    int32_t orig_w, orig_h, orig_channels;   /* geometry of `original` when orig_is_input == 0 */
    int32_t magnified;                       /* the magnification stage produced a frame */
} mc_chain_info;
mc_status mc_chain_process(mc_handle* h, const uint8_t* in, int width, int height, int channels, size_t in_step,
                           const mc_params* p, int grayscale, uint8_t* out, size_t out_bytes, uint8_t* original,
                           size_t original_bytes, mc_chain_info* info);

/* Pipelined host path: mc_submit enqueues H2D + kernels + D2H for one frame asynchronously on
 * three streams (copy-in, compute, copy-out) and returns; mc_collect waits for the OLDEST
 * outstanding frame (strict FIFO — frame order is preserved, as ProcessingChain.hpp:18-20 needs)
 * and reports whether `out` of that submit was written.  `in`/`out` that are pinned host memory
 * (cudaHostAlloc / cudaHostRegister / mc_host_alloc) are copied straight to/from HBM; pageable
 * buffers go through the handle's pinned staging slots.  At most mc_pipeline_depth() frames may
 * be in flight; `in` and `out` must stay valid until the frame is collected. */
int mc_pipeline_depth(mc_handle* h);
mc_status mc_submit(mc_handle* h, const uint8_t* in, int width, int height, int channels,
                    size_t in_step, const mc_params* p, uint8_t* out, size_t out_step);
mc_status mc_collect(mc_handle* h, int* produced);

/* Pinned host memory for frames (so FramePool buffers can be DMA'd directly). */
void* mc_host_alloc(size_t bytes);
void mc_host_free(void* p);

/* The handle's cudaStream_t (as void*) so callers can order their own work / record events on it. */
void* mc_stream(mc_handle* h);

/* Options (call before the first frame or after mc_reset):
 *   "faithful_level0" (default 0): also run the level-0 band + IIR state update that the reference
 *        performs although the gain loop multiplies that band by 0 (MagnifyCore.hpp:130-131); needed
 *        only to compare level-0 state planes with the oracle.
 *   "pipeline_depth"  (default 3)
 *   "keep_float_output" (default 0): keep the pre-quantisation float image (tests)
 *   "profile_kernels" (default 0): see mc_profile_read
 *   "use_tma" (default 1): stage the fused level kernel's tiles with TMA (cp.async.bulk.tensor); 0 selects
 *        the 128-bit LDG staging path (same results; kept for A/B measurements)
 *   "prefetch_state" (default 1; needs use_tma): the fused level kernel requests the tile's two state planes as TMA
 *        bulk copies at kernel entry, together with its input window, instead of loading them in its last phase
 *        (same results; measured on B200: level[1] 235 -> 205 us per 32-lane launch; 0 kept for A/B measurements)
 *   "lane_groups" (default 0 = automatic: min(2, lanes / 8), at least 1): Laplace — the lanes of the handle run as
 *        that many concurrent launch chains on separate CUDA streams, forked from and joined into mc_stream(); the
 *        L1-bound ingest, issue-bound egress and HBM-bound level kernels of different groups then share the SMs
 *        instead of running back to back (same results; profile_kernels forces 1)
 *   "egress_strip" (default 20): Laplace egress runs as the shuffle strip kernel (one warp per 128-column strip,
 *        sliding windows in per-lane shared-memory rings; the value 16 / 20 / 24 picks the register cap = resident
 *        warps per SM); 0 selects the shared-memory tile kernel (bit-identical results; kept for A/B)
 *   "ingest_warps" (default 1): warps per CTA (1, 2 or 4) of the fused BGR->Lab ingest kernel (same results; A/B)
 *   "analysis_only" (default 0): Laplace and Phase — frames after the first update the temporal state (EMA planes;
 *        Riesz pyramids, phase accumulators and Butterworth registers) but skip synthesis and egress and report
 *        *produced = 0; the cheap first pass of temporal sharding
 *   "band_from_state" (default 1): Laplace synthesis rebuilds each amplified band gain*(hi-lo) from the two
 *        state planes instead of reading a band plane stored by the level kernel (same results; takes 4 B/px off
 *        the level kernel's interface and adds them to the collapse / egress kernels; measured on B200 together
 *        with prefetch_state: level[1] 205 -> 177 us, egress +8 us, step -1.6 %; 0 kept for A/B measurements) */
mc_status mc_set_option(mc_handle* h, const char* key, int value);

/* Test-only access to temporal state planes as dense f32 [lanes][channels][rows][cols].
 * Names: Laplace "lowpassHi" / "lowpassLo" (MotionState, MagnifyCore.hpp:24-29), level 0..levels (levels 0 and
 *   `levels` exist only with option faithful_level0);
 * Phase, per band level 0..levels-2, one channel: "old.lowpass", "old.rx", "old.ry" (RieszState::old),
 *   "phase.c", "phase.s" (itsPhase — the two filters' copies are identical), "lo.r0.c", "lo.r0.s", "lo.r1.c",
 *   "lo.r1.s", "hi.r0.c", "hi.r0.s", "hi.r1.c", "hi.r1.s" (itsRegister0/1 of the low / high cutoff filters,
 *   TemporalFilter.cpp:299-317);
 * Color keeps its rolling window in a device ring buffer that is not exposed.
 * mc_state_dims reports rows/cols/channels for a name+level (0 rows if absent). */
mc_status mc_state_dims(mc_handle* h, const char* name, int level, int* rows, int* cols, int* channels);
mc_status mc_get_state(mc_handle* h, const char* name, int level, float* dst, size_t dst_floats);
mc_status mc_set_state(mc_handle* h, const char* name, int level, const float* src, size_t src_floats);

/* Debug tap: last frame's pre-quantisation BGR/gray float output in [0,1] ([lanes][rows][cols][C]);
 * enabled by option "keep_float_output" = 1. */
mc_status mc_get_float_output(mc_handle* h, float* dst, size_t dst_floats);

/* Number of kernel launches this handle has issued (for bench.py's gpu_launches). */
uint64_t mc_launch_count(mc_handle* h);

/* Per-kernel device timing for bench.py's roofline: with option "profile_kernels" = 1 every launch
 * is bracketed by CUDA events on the handle's stream.  mc_profile_read synchronises, drains them and
 * writes text lines "kernel level launches total_ms\n" (NUL-terminated) into buf. */
mc_status mc_profile_read(mc_handle* h, char* buf, size_t cap);

/* Last error text for this handle (or for a failed mc_create when h == NULL); never NULL. */
const char* mc_last_error(mc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MAGCORE_B200_H */
