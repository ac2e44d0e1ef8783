// Exhaustive check of obvhs_cwbvh_exponent (product header) against the platform libm the reference would call
// (f32::log2 / f32::exp2 -> glibc log2f / exp2f), reference src/cwbvh/bvh2_to_cwbvh.rs:85-99.
// Covers every f32 in [lo_bits, +inf]. Prints "OK <count>" or the first mismatches. Built and run by
// tests/test_exponent_exhaustive.py.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../obvhs_b200/csrc/cwbvh_exponent.h"

int main(int argc, char** argv) {
    uint32_t lo = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 0) : 0x19000000u;  // ~6.6e-24 < 1e-20/255
    uint32_t hi = 0x7f800000u;                                                 // +inf inclusive
    long long bad = 0, count = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad, count)
    for (long long b = lo; b <= (long long)hi; b++) {
        uint32_t bits = (uint32_t)b;
        float v;
        memcpy(&v, &bits, 4);
        float e = exp2f(ceilf(log2f(v)));
        float rcp = 1.0f / e;
        uint32_t eb;
        memcpy(&eb, &e, 4);
        uint8_t want = (uint8_t)(eb >> 23);
        float got_rcp;
        uint8_t got = obvhs_cwbvh_exponent(v, &got_rcp);
        uint32_t r0, r1;
        memcpy(&r0, &rcp, 4);
        memcpy(&r1, &got_rcp, 4);
        if (got != want || r0 != r1) {
            if (bad < 10) fprintf(stderr, "mismatch v=%a bits=%08x want e=%u rcp=%a got e=%u rcp=%a\n", v, bits, want, rcp, got, got_rcp);
            bad++;
        }
        count++;
    }
    if (bad) {
        printf("FAIL %lld of %lld\n", bad, count);
        return 1;
    }
    printf("OK %lld\n", count);
    return 0;
}
