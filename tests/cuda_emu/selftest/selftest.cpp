// Self-test of the CUDA-on-CPU emulation (tests/cuda_emu): small kernels whose results are known in closed form, run
// under whatever CUDA_EMU_ORDER / CUDA_EMU_ASYNC the environment selects.  Exit code 0 = all checks passed.
#include <cuda.h>

#include <cstdio>
#include <numeric>
#include <vector>

static int g_fail = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
    } while (0)

// 1. shuffles: idx / up / down / xor, several rounds, lanes exiting right after their last shuffle
__global__ void k_shfl(int* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int v = lane + 100 * warp;
    int acc = 0;
    for (int r = 0; r < 5; ++r) {
        acc += __shfl_up_sync(0xffffffffu, v, 1) - __shfl_down_sync(0xffffffffu, v, 2) + __shfl_xor_sync(0xffffffffu, v, 5) +
               __shfl_sync(0xffffffffu, v, 7);
        v += 3;
    }
    out[threadIdx.x] = acc;
}
static int ref_shfl(int lane, int warp) {
    int acc = 0;
    for (int r = 0; r < 5; ++r) {
        auto val = [&](int l) { return l + 100 * warp + 3 * r; };
        acc += val(lane >= 1 ? lane - 1 : lane) - val(lane + 2 <= 31 ? lane + 2 : lane) + val(lane ^ 5) + val(7);
    }
    return acc;
}

// 2. barriers with threads that exit early, shared memory written by one half and read by the other
__global__ void k_bar(const int* in, int* out) {
    __shared__ int s[256];
    if (threadIdx.x >= 200) return;               // exited threads do not take part in later barriers
    s[threadIdx.x] = in[threadIdx.x] * 2;
    __syncthreads();
    const int v = s[199 - threadIdx.x];
    __syncthreads();
    s[threadIdx.x] = v + 1;
    __syncthreads();
    out[threadIdx.x] = s[(threadIdx.x + 1) % 200];
}

// 3. TMA: two tiled loads (one partly out of bounds) on ONE mbarrier, waited for by every thread
__global__ void k_tma(const __grid_constant__ CUtensorMap tm, float* out, int x, int y) {
    __shared__ __align__(128) float a[8][16];
    __shared__ __align__(128) float b[8][16];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) cuda_emu::mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        cuda_emu::mbar_expect_tx(&bar, 2 * 8 * 16 * sizeof(float));
        const int c0[3] = {x, y, 0}, c1[3] = {-3, -2, 1};
        cuda_emu::tma_load(&a[0][0], &tm, c0, &bar);
        cuda_emu::tma_load(&b[0][0], &tm, c1, &bar);
    }
    cuda_emu::mbar_wait(&bar, 0);
    const int r = threadIdx.x / 16, c = threadIdx.x % 16;
    if (r < 8) { out[r * 16 + c] = a[r][c]; out[128 + r * 16 + c] = b[r][c]; }
}

// 4. atomics
__global__ void k_atomic(unsigned* mn, unsigned* mx, int* sum) {
    const unsigned v = (threadIdx.x * 2654435761u) >> 8;
    atomicMin(mn, v); atomicMax(mx, v); atomicAdd(sum, (int)threadIdx.x);
}

// 5. streams: producer on stream A, consumer on stream B behind an event
__global__ void k_fill(int* p, int n, int v) { for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = v + i; }
__global__ void k_scale(const int* in, int* out, int n) { for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = in[i] * 3; }

int main() {
    {   // shuffles
        int* d; cudaMalloc((void**)&d, 96 * sizeof(int));
        cuda_emu::Launcher(dim3(1), dim3(96)).run("k_shfl", (k_shfl), d);
        cudaDeviceSynchronize();
        for (int t = 0; t < 96; ++t) CHECK(d[t] == ref_shfl(t & 31, t >> 5));
        cudaFree(d);
    }
    {   // barriers
        std::vector<int> in(256), out(256, -1);
        std::iota(in.begin(), in.end(), 5);
        cuda_emu::Launcher(dim3(3), dim3(256)).run("k_bar", (k_bar), (const int*)in.data(), out.data());
        cudaDeviceSynchronize();
        for (int t = 0; t < 200; ++t) CHECK(out[t] == in[199 - (t + 1) % 200] * 2 + 1);
        CHECK(out[200] == -1);
    }
    {   // TMA + mbarrier
        const int W = 40, H = 24, P = 2, pitch = 48;
        float* g; cudaMalloc((void**)&g, sizeof(float) * pitch * H * P);
        for (int i = 0; i < pitch * H * P; ++i) g[i] = (float)i;
        void* fn = nullptr;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault);
        typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        CUtensorMap tm;
        const cuuint64_t dims[3] = {W, H, P}, strides[2] = {pitch * 4, (cuuint64_t)pitch * H * 4};
        const cuuint32_t box[3] = {16, 8, 1}, es[3] = {1, 1, 1};
        CHECK(((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS);
        const cuuint32_t badbox[3] = {15, 8, 1};   // 60 bytes per row: not a multiple of 16 -> must be rejected
        CUtensorMap bad;
        CHECK(((Enc)fn)(&bad, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, dims, strides, badbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS);
        float* out; cudaMalloc((void**)&out, 256 * sizeof(float));
        cuda_emu::Launcher(dim3(1), dim3(160)).run("k_tma", (k_tma), tm, out, 30, 20);
        cudaDeviceSynchronize();
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 16; ++c) {
                const int gx = 30 + c, gy = 20 + r;
                const float want = (gx < W && gy < H) ? (float)(gy * pitch + gx) : 0.0f;
                CHECK(out[r * 16 + c] == want);
                const int hx = -3 + c, hy = -2 + r;
                const float want2 = (hx >= 0 && hy >= 0) ? (float)(pitch * H + hy * pitch + hx) : 0.0f;
                CHECK(out[128 + r * 16 + c] == want2);
            }
        cudaFree(out); cudaFree(g);
    }
    {   // atomics
        unsigned mn = 0xffffffffu, mx = 0; int sum = 0;
        cuda_emu::Launcher(dim3(2), dim3(128)).run("k_atomic", (k_atomic), &mn, &mx, &sum);
        cudaDeviceSynchronize();
        unsigned rmn = 0xffffffffu, rmx = 0;
        for (unsigned t = 0; t < 128; ++t) { const unsigned v = (t * 2654435761u) >> 8; rmn = v < rmn ? v : rmn; rmx = v > rmx ? v : rmx; }
        CHECK(mn == rmn && mx == rmx && sum == 2 * (127 * 128 / 2));
    }
    {   // streams + events: correct with the wait; (with CUDA_EMU_ASYNC=1) wrong for some schedule without it
        cudaStream_t sa, sb; cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
        cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        const int n = 64;
        int *a, *b; cudaMalloc((void**)&a, n * sizeof(int)); cudaMalloc((void**)&b, n * sizeof(int));
        int wrong_without_wait = 0;
        for (int trial = 0; trial < 40; ++trial) {
            for (int with_wait = 1; with_wait >= 0; --with_wait) {
                cudaMemsetAsync(a, 0, n * sizeof(int), sa);
                cudaStreamSynchronize(sa);
                cuda_emu::Launcher(dim3(1), dim3(32), 0, sa).run("k_fill", (k_fill), a, n, trial);
                cudaEventRecord(ev, sa);
                if (with_wait) cudaStreamWaitEvent(sb, ev, 0);
                cuda_emu::Launcher(dim3(1), dim3(32), 0, sb).run("k_scale", (k_scale), (const int*)a, b, n);
                cudaStreamSynchronize(sa); cudaStreamSynchronize(sb);
                bool ok = true;
                for (int i = 0; i < n; ++i) ok = ok && b[i] == 3 * (trial + i);
                if (with_wait) CHECK(ok); else wrong_without_wait += !ok;
            }
        }
        const char* as = std::getenv("CUDA_EMU_ASYNC");
        if (as && as[0] == '1') CHECK(wrong_without_wait > 0);   // the missing dependency must be observable
        else CHECK(wrong_without_wait == 0);                     // immediate mode = program order
        cudaEventDestroy(ev); cudaFree(a); cudaFree(b); cudaStreamDestroy(sa); cudaStreamDestroy(sb);
    }
    std::printf(g_fail ? "selftest: %d check(s) failed\n" : "selftest: ok\n", g_fail);
    return g_fail ? 1 : 0;
}
