// cuda_emu — TEST INFRASTRUCTURE ONLY.  Driver-API slice used by the kernels: CUtensorMap + its encoder types,
// and the emulated bulk-tensor (TMA) load / mbarrier primitives the level kernel is built on.
#pragma once
#include <cuda_runtime.h>

typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUresult { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0, CU_TENSOR_MAP_DATA_TYPE_UINT16 = 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_64B = 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

// 128 opaque bytes on the device; here the emulated encoder keeps the description in the clear
struct alignas(64) CUtensorMap {
    union {
        unsigned char opaque[128];
        struct {
            const unsigned char* base;
            uint64_t dims[5], strides[4];   // strides in bytes for dims 1..rank-1
            uint32_t box[5], elem_bytes, rank;
        } emu;
    };
};

namespace cuda_emu {
// tiled load of tm.box at coordinates c[] into dst (packed, x fastest); out-of-bounds elements are zero-filled;
// completes the mbarrier `bar` (one phase).
void tma_load(void* dst, const CUtensorMap* tm, const int* coords, uint64_t* bar);
void tma_store(const void* src, const CUtensorMap* tm, const int* coords);
// mbarrier with transaction counting: a phase completes when the expected arrivals have arrived AND the expected
// bytes have landed (mbarrier.arrive.expect_tx / complete_tx).  *bar holds the number of completed phases.
void mbar_init(uint64_t* bar, unsigned arrivals);
void mbar_expect_tx(uint64_t* bar, unsigned bytes);        // one arrival + `bytes` expected
void mbar_complete_tx(uint64_t* bar, unsigned bytes);      // what a finished bulk copy does
inline void mbar_wait(uint64_t* bar, unsigned parity) {
    while (((*bar) & 1u) == parity) yield_spin();
}
}  // namespace cuda_emu
