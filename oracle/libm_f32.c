/* TEST INFRASTRUCTURE ONLY (oracle/).  The reference evaluates acosf / cosf / sinf from the C library element by
 * element (src/processing/magnification/RieszPyramid.cpp:8-38: arcCos, cosSin).  numpy's float32 ufuncs use their own
 * SIMD kernels, which differ from glibc in the last ulp on ~1/3 of the samples, so the oracle calls the same libm
 * functions through this file (built by oracle/build_ref.py / livim_oracle._libm()). */
#include <math.h>
#include <stddef.h>

/* arcCos, RieszPyramid.cpp:8-23: out-of-range arguments return -1.0 / +1.0 (radians), not pi / 0 */
void livim_arccos_f32(const float* x, float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (x[i] < -1.0) out[i] = -1.0;
        else if (x[i] > 1.0) out[i] = 1.0;
        else out[i] = acosf(x[i]);
    }
}

/* cosSin, RieszPyramid.cpp:25-38 */
void livim_cossin_f32(const float* x, float* c, float* s, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        c[i] = cosf(x[i]);
        s[i] = sinf(x[i]);
    }
}
