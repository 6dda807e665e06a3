/* oracle/bvh_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See bvh_oracle.h.
 * Instantiates bvh_oracle_impl.inc for Node<float,3> and Node<double,3>.
 * Compile: gcc -std=c11 -O3 -march=x86-64-v3 -ffp-contract=off -fPIC -shared (oracle/Makefile). */
#define _GNU_SOURCE
#include "bvh_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* ---- float ---- */
#define T float
#define UT uint32_t
#define S(x) CAT(x, 3f)
#define T_MAX FLT_MAX
#define T_EPS FLT_EPSILON
#define FABS fabsf
#define COPYSIGN copysignf
#define FMA fmaf
#define ISFINITE isfinite
#include "bvh_oracle_impl.inc"
#undef T
#undef UT
#undef S
#undef T_MAX
#undef T_EPS
#undef FABS
#undef COPYSIGN
#undef FMA
#undef ISFINITE

/* ---- double ---- */
#define T double
#define UT uint64_t
#define S(x) CAT(x, 3d)
#define T_MAX DBL_MAX
#define T_EPS DBL_EPSILON
#define FABS fabs
#define COPYSIGN copysign
#define FMA fma
#define ISFINITE isfinite
#include "bvh_oracle_impl.inc"
#undef T
#undef UT
#undef S

/* utils.h:103-120: spread the low third of the bits so that two zero bits separate them */
static uint32_t split_bits32(uint32_t x) {
    uint32_t mask = 0xFFFFFFFFu >> 16;
    x &= mask;
    for (uint32_t n = 16; n > 1; n >>= 1) {      /* i = log_bits-1 .. 1, n = 1 << i */
        mask = (mask | (mask << n)) & ~(mask << (n / 2));
        x = (x | (x << n)) & mask;
    }
    return x;
}
static uint64_t split_bits64(uint64_t x) {
    uint64_t mask = 0xFFFFFFFFFFFFFFFFull >> 32;
    x &= mask;
    for (uint64_t n = 32; n > 1; n >>= 1) {
        mask = (mask | (mask << n)) & ~(mask << (n / 2));
        x = (x | (x << n)) & mask;
    }
    return x;
}
uint32_t orc_morton_encode32(uint32_t x, uint32_t y, uint32_t z) {
    return split_bits32(x) | (split_bits32(y) << 1) | (split_bits32(z) << 2);
}
uint64_t orc_morton_encode64(uint64_t x, uint64_t y, uint64_t z) {
    return split_bits64(x) | (split_bits64(y) << 1) | (split_bits64(z) << 2);
}

int orc_fast_mul_add_is_fma(void) {
#ifdef FP_FAST_FMAF
    return 1;
#else
    return 0;
#endif
}
