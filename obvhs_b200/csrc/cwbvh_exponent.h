// cwbvh_exponent.h -- the CwBvhNode exponent bytes, bit-exact with the reference on linux-gnu.
//
// The reference computes (src/cwbvh/bvh2_to_cwbvh.rs:85-88), per axis,
//     e = exp2(ceil(log2(max(extent, 1e-20) * (1/255))))
// through f32::log2 / f32::exp2, i.e. glibc's log2f / exp2f. exp2f of an integer is an exact power of two, so the
// only subtle step is ceil(log2f(v)): for v = 2^k * m with 1 < m < 2 the true logarithm is k + log2(m) > k, but
// when m is within a few ulps of 1 and |k| >= 4 the float result ROUNDS to exactly k, and the reference then picks
// the exponent one binade too small. We reproduce that rounding with f64 arithmetic:
//     fl32(k + log2(m)) > k  <=>  ceil = k + 1.
// The closest any candidate m comes to the rounding boundary is ~0.8% of half a float ulp of k, orders of
// magnitude above the error of an f64 log2 or of glibc's own f64 polynomial, so the decision is unambiguous.
// tests/test_exponent_exhaustive.py checks this function against glibc's log2f/exp2f for every positive float.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OBVHS_HD __host__ __device__
#else
#define OBVHS_HD
#endif

// v > 0 (normal, possibly +inf). Returns the biased exponent byte of e (f32 bits >> 23, truncated to u8 exactly like
// the reference's `as u8`), and *rcp_e = 1.0 / e.
OBVHS_HD inline uint8_t obvhs_cwbvh_exponent(float v, float* rcp_e) {
    uint32_t bits;
#if defined(__CUDA_ARCH__)
    bits = __float_as_uint(v);
#else
    memcpy(&bits, &v, 4);
#endif
    int biased = (int)((bits >> 23) & 0xffu);
    uint32_t mant = bits & 0x7fffffu;
    if (biased == 0xff) {  // +inf (or NaN): log2f -> inf, exp2f(inf) = inf, bits >> 23 = 255, 1/inf = 0
        *rcp_e = 0.0f;
        return 255;
    }
    int k = biased - 127;
    int kk = k;
    if (mant != 0) {
        uint32_t mb = 0x3f800000u | mant;
        float m;
#if defined(__CUDA_ARCH__)
        m = __uint_as_float(mb);
#else
        memcpy(&m, &mb, 4);
#endif
        double y = (double)k + log2((double)m);
        float yf = (float)y;
        kk = (yf > (float)k) ? k + 1 : k;
    }
    if (kk >= 128) {  // exp2f overflows to +inf
        *rcp_e = 0.0f;
        return 255;
    }
    // 2^kk for kk in [-126, 127]; smaller kk cannot occur because extent >= 1e-20
    uint32_t rb = (uint32_t)(127 - kk) << 23;
    float r;
#if defined(__CUDA_ARCH__)
    r = kk == 127 ? 5.877471754111438e-39f : __uint_as_float(rb);
#else
    if (kk == 127) r = 5.877471754111438e-39f;  // 2^-127 (denormal), what 1.0f / 2^127 evaluates to
    else memcpy(&r, &rb, 4);
#endif
    *rcp_e = r;
    return (uint8_t)(kk + 127);
}
