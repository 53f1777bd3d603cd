// cuda_emu — TEST INFRASTRUCTURE ONLY.  cuFFT slice used by Color mode: batched strided 1-D R2C / C2R, evaluated
// as a direct O(n^2) DFT in double precision (n <= 64 on this path).
#pragma once
#include <cuda_runtime.h>

typedef int cufftHandle;
typedef float cufftReal;
typedef float2 cufftComplex;
enum cufftResult { CUFFT_SUCCESS = 0, CUFFT_INVALID_PLAN = 1, CUFFT_ALLOC_FAILED = 2, CUFFT_INVALID_VALUE = 4, CUFFT_INTERNAL_ERROR = 5 };
enum cufftType { CUFFT_R2C = 0x2a, CUFFT_C2R = 0x2c, CUFFT_C2C = 0x29 };

cufftResult cufftPlanMany(cufftHandle* plan, int rank, int* n, int* inembed, int istride, int idist, int* onembed, int ostride,
                          int odist, cufftType type, int batch);
// plan objects with a caller-owned work area (the product shares one work area between its cached plans)
cufftResult cufftCreate(cufftHandle* plan);
cufftResult cufftSetAutoAllocation(cufftHandle plan, int autoAllocate);
cufftResult cufftMakePlanMany(cufftHandle plan, int rank, int* n, int* inembed, int istride, int idist, int* onembed, int ostride,
                              int odist, cufftType type, int batch, size_t* workSize);
cufftResult cufftSetWorkArea(cufftHandle plan, void* workArea);
cufftResult cufftSetStream(cufftHandle plan, cudaStream_t s);
cufftResult cufftExecR2C(cufftHandle plan, cufftReal* in, cufftComplex* out);
cufftResult cufftExecC2R(cufftHandle plan, cufftComplex* in, cufftReal* out);
cufftResult cufftDestroy(cufftHandle plan);
