// fastmath.cuh -- FP64 exp / sincos specialised to the argument ranges of this path.
//
// The wACSF kernels spend most of their FP64 work in exp(-alpha*s) (s >= 0) and in
// cos/sin(pi~ r/Rc) with r <= Rc.  The CUDA library versions carry range checks,
// huge-argument reduction and denormal handling that are dead weight here (ncu:
// ~50 SASS instructions per exp, ~70 per sincos).  These versions keep full double
// accuracy (<= ~2 ulp; tests/test_fastmath.py checks them against libm on the host
// with the same source) on the restricted domains:
//   exp_neg(x)      -700 <= x <= 0  (callers clamp the argument when it can be lower)
//   sincos_0pi(y)   0 <= y <= 3.3
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define GAPCU_HD __host__ __device__ __forceinline__
#else
#define GAPCU_HD static inline
#endif

namespace gapcu {

// 2^(j/32), j = 0..31; the kernels copy it to shared memory (per-lane index).
static inline void fill_exp2_table(double *t) {
    for (int j = 0; j < 32; j++) t[j] = exp2(j / 32.0);
}

// Polynomial / reduction constants.  On the device they live in __constant__ memory so
// that DFMA takes them straight from the constant bank; as literals every one of them
// costs two extra (U)MOV instructions per use (seen in SASS and in the ncu source page).
#define GAPCU_KC_LIST                                                                        \
    46.16624130844683,        /* 0  32/ln2                        */                         \
    -0.021660849392446835,    /* 1  -ln2/32 high 36 bits          */                         \
    -5.145609244655338e-14,   /* 2  -ln2/32 low                   */                         \
    0.001388888888888889,     /* 3  1/6!                          */                         \
    0.008333333333333333,     /* 4  1/5!                          */                         \
    0.041666666666666664,     /* 5  1/4!                          */                         \
    0.16666666666666666,      /* 6  1/3!                          */                         \
    0.6366197723675814,       /* 7  2/pi                          */                         \
    -1.5707963267948912,      /* 8  -pi/2 high                    */                         \
    -5.390302858158119e-15,   /* 9  -pi/2 low                     */                         \
    2.8114572543455206e-15, -7.647163731819816e-13, 1.6059043836821613e-10, /* 10-12 sin */  \
    -2.505210838544172e-08, 2.7557319223985893e-06, -0.0001984126984126984, /* 13-15     */  \
    0.008333333333333333, -0.16666666666666666,                             /* 16-17     */  \
    4.779477332387385e-14, -1.1470745597729725e-11, 2.08767569878681e-09,   /* 18-20 cos */  \
    -2.755731922398589e-07, 2.48015873015873e-05, -0.001388888888888889,    /* 21-23     */  \
    0.041666666666666664                                                    /* 24        */

#if defined(__CUDACC__)
__constant__ double gapcu_kc_dev[] = {GAPCU_KC_LIST};
#endif
static const double gapcu_kc_host[] = {GAPCU_KC_LIST};
#if defined(__CUDA_ARCH__)
#define KC(i) gapcu_kc_dev[i]
#else
#define KC(i) gapcu_kc_host[i]
#endif

// exp(x) for -700 <= x <= 0 (callers clamp).  x = (32 m + j) ln2/32 + r, |r| <= ln2/64:
// exp(x) = 2^m * T[j] * (1 + r + r^2/2! + ... + r^6/6!)        (next term < 4e-18)
GAPCU_HD double exp_neg(double x, const double *T32) {
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: rounds to nearest integer
    const double kd = fma(x, KC(0), MAGIC);
    const double kf = kd - MAGIC;
    long long kbits;
    memcpy(&kbits, &kd, sizeof kbits);
    const int ki = (int)(unsigned int)kbits;             // low word of kd holds the integer (two's complement)
    double r = fma(kf, KC(1), x);                        // k*hi is exact
    r = fma(kf, KC(2), r);
    double p = fma(r, KC(3), KC(4));
    p = fma(p, r, KC(5));
    p = fma(p, r, KC(6));
    p = fma(p, r, 0.5);
    p = fma(p, r * r, r);                                // expm1(r)
    const double t = T32[ki & 31];
    double y = fma(t, p, t);
    unsigned long long yb;
    memcpy(&yb, &y, sizeof yb);
    yb += (unsigned long long)(unsigned int)((ki >> 5) << 20) << 32;  // * 2^m on the high word (y in [1,2.1), m >= -1010)
    memcpy(&y, &yb, sizeof y);
    return y;
}

// sin(y), cos(y) for 0 <= y <= 3.3: quadrant q in {0,1,2}, t = y - q*pi/2 in [-pi/4, pi/4]
GAPCU_HD void sincos_0pi(double y, double *sn, double *cs) {
    const double q = rint(y * KC(7));
    double t = fma(q, KC(8), y);                         // q*hi exact
    t = fma(q, KC(9), t);
    const double z = t * t, z2 = z * z, z4 = z2 * z2;
    // Estrin evaluation: four independent pairs per polynomial instead of one 8-long chain
    const double s01 = fma(z, KC(16), KC(17)), s23 = fma(z, KC(14), KC(15));
    const double s45 = fma(z, KC(12), KC(13)), s67 = fma(z, KC(10), KC(11));
    double s = fma(z4, fma(z2, s67, s45), fma(z2, s23, s01));
    s = fma(s * z, t, t);                                // sin(t)
    const double c01 = fma(z, -0.5, 1.0), c23 = fma(z, KC(23), KC(24));
    const double c45 = fma(z, KC(21), KC(22)), c67 = fma(z, KC(19), KC(20));
    const double c03 = fma(z2, c23, c01), c47 = fma(z2, c67, c45);
    double c = fma(z4, fma(z4, KC(18), c47), c03);       // cos(t): degree 8 in z
    // q = 0: (s, c); q = 1: (c, -s); q = 2: (-s, -c)
    const bool odd = (q == 1.0);
    const bool two = (q == 2.0);
    const double ss = two ? -s : s, cc = two ? -c : c;
    *sn = odd ? c : ss;
    *cs = odd ? -s : cc;
}

#if defined(__CUDACC__)
// 1/sqrt(x) for normal positive x: hardware seed (rsqrt.approx.ftz.f64, ~2^-22) refined by two
// Newton steps (y <- y + y*(1 - x y^2)/2): ~1 ulp, 10 instructions instead of the ~35 of the
// library routine, which also handles zero / denormal / infinite arguments.
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);
    y = fma(y * 0.5, e, y);
    e = fma(-x * y, y, 1.0);
    y = fma(y * 0.5, e, y);
    return y;
}
#endif

}  // namespace gapcu
