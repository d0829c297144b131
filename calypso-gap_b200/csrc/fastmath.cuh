// fastmath.cuh -- FP64 exp / sincos specialised to the argument ranges of this path.
//
// The wACSF kernels spend most of their FP64 work in exp(-alpha*s) (s >= 0) and in
// cos/sin(pi~ r/Rc) with r <= Rc.  The CUDA library versions carry range checks,
// huge-argument reduction and denormal handling that are dead weight here (ncu:
// ~50 SASS instructions per exp, ~70 per sincos).  These versions keep full double
// accuracy (<= ~2 ulp; tests/test_boundary_cpu.py::test_fastmath_host_versions checks
// them against libm on the host with the same source) on the restricted domains:
//   exp_neg(x)      x <= 0; arguments below about -708 return a value <= 2^-1021 instead
//                   of a denormal or zero (one integer max on the exponent, no FP clamp)
//   sincos_0pi(y)   0 <= y <= 3.3: no quadrant logic, both functions are polynomials in
//                   t = y - pi/2 (sin y = cos t, cos y = -sin t)
//   sincos_tab(y)   same range, through a 54-entry table of (cos, sin)(k/16) and the angle-addition
//                   formulas with |delta| <= 1/32: 18 FP64 operations and 6 constants instead of 31
//                   and 20 (the hot loops' version; the table is filled with sincos_0pi)
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
    -1.5707963267948966,      /* 7  -pi/2 high                    */                         \
    -6.123233995736766e-17,   /* 8  -pi/2 low                     */                         \
    /* 9-18: sin t = t + t^3 (S1 + S2 z + ... + S10 z^9), z = t^2 (Taylor; next term < 2e-18 on |t| <= 1.73) */ \
    -0.16666666666666666, 0.008333333333333333, -0.0001984126984126984, 2.7557319223985893e-06,   \
    -2.505210838544172e-08, 1.6059043836821613e-10, -7.647163731819816e-13, 2.8114572543455206e-15, \
    -8.22063524662433e-18, 1.9572941063391263e-20,                                                 \
    /* 19-28: cos t = 1 - z/2 + z^2 (C2 + C3 z + ... + C11 z^9) */                                \
    0.041666666666666664, -0.001388888888888889, 2.48015873015873e-05, -2.755731922398589e-07,     \
    2.08767569878681e-09, -1.1470745597729725e-11, 4.779477332387385e-14, -1.5619206968586225e-16, \
    4.110317623312165e-19, -8.896791392450574e-22,                                                  \
    /* 29-31: cos d - 1 = z (-1/2 + z (C2 + z (C3 + z C4))), z = d^2, |d| <= 1/32 (next term 2e-22) */ \
    0.041666666666666664, -0.001388888888888889, 2.48015873015873e-05,                              \
    /* 32-34: sin d = d + d z (S1 + z (S2 + z S3))          (next term 3e-18 relative) */          \
    -0.16666666666666666, 0.008333333333333333, -0.0001984126984126984

#if defined(__CUDACC__)
__constant__ double gapcu_kc_dev[] = {GAPCU_KC_LIST};
#endif
static const double gapcu_kc_host[] = {GAPCU_KC_LIST};
#if defined(__CUDA_ARCH__)
#define KC(i) gapcu_kc_dev[i]
#else
#define KC(i) gapcu_kc_host[i]
#endif

// exp(x) for x <= 0.  x = (32 m + j) ln2/32 + r, |r| <= ln2/64:
// exp(x) = 2^m * T[j] * (1 + r + r^2/2! + ... + r^6/6!)        (next term < 4e-18)
// Two integer clamps replace the FP one: the high word of x is capped at that of -1e7 (as
// unsigned integers, larger means more negative), and m is kept >= -1021 so that the
// exponent field cannot wrap; results that should underflow come out as <= 2^-1021.
GAPCU_HD double exp_neg(double x, const double *T32) {
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: rounds to nearest integer
    unsigned long long xb;
    memcpy(&xb, &x, sizeof xb);
    unsigned int xh = (unsigned int)(xb >> 32);
    xh = xh < 0xc16312d0u ? xh : 0xc16312d0u;            // x >= -1e7 (one integer min)
    xb = ((unsigned long long)xh << 32) | (xb & 0xffffffffull);
    memcpy(&x, &xb, sizeof x);
    const double kd = fma(x, KC(0), MAGIC);
    const double kf = kd - MAGIC;
    long long kbits;
    memcpy(&kbits, &kd, sizeof kbits);
    const int ki = (int)(unsigned int)kbits;             // low word of kd holds the integer (two's complement)
    double r = fma(kf, KC(1), x);                        // k*hi is exact for |k| < 2^17, accurate enough beyond
    r = fma(kf, KC(2), r);
    double p = fma(r, KC(3), KC(4));
    p = fma(p, r, KC(5));
    p = fma(p, r, KC(6));
    p = fma(p, r, 0.5);
    p = fma(p, r * r, r);                                // expm1(r)
    const double t = T32[ki & 31];
    double y = fma(t, p, t);
    int m = ki >> 5;
    m = m > -1021 ? m : -1021;
    unsigned long long yb;
    memcpy(&yb, &y, sizeof yb);
    yb += (unsigned long long)(unsigned int)(m << 20) << 32;  // * 2^m on the high word (y in [1,2.1))
    memcpy(&y, &yb, sizeof y);
    return y;
}

// sin(y), cos(y) for 0 <= y <= 3.3 through t = y - pi/2 in [-1.571, 1.73]
GAPCU_HD void sincos_0pi(double y, double *sn, double *cs) {
    const double t = (y + KC(7)) + KC(8);
    const double z = t * t, z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    // Estrin evaluation: independent pairs instead of one 10-long chain per polynomial
    const double s01 = fma(z, KC(10), KC(9)), s23 = fma(z, KC(12), KC(11)), s45 = fma(z, KC(14), KC(13));
    const double s67 = fma(z, KC(16), KC(15)), s89 = fma(z, KC(18), KC(17));
    double s = fma(z4, fma(z2, s67, s45), fma(z2, s23, s01));
    s = fma(z8, s89, s);
    s = fma(t * z, s, t);                                // sin(t)
    const double c01 = fma(z, KC(20), KC(19)), c23 = fma(z, KC(22), KC(21)), c45 = fma(z, KC(24), KC(23));
    const double c67 = fma(z, KC(26), KC(25)), c89 = fma(z, KC(28), KC(27));
    double c = fma(z4, fma(z2, c67, c45), fma(z2, c23, c01));
    c = fma(z8, c89, c);
    c = fma(z2, c, fma(z, -0.5, 1.0));                   // cos(t)
    *sn = c;
    *cs = -s;
}

// sin(y), cos(y) for 0 <= y <= 3.3 with a table T[k] = (cos(k/16), sin(k/16)), k = 0..53 (SINCOS_TAB_N
// entries, filled by fill_sincos_table or, on the device, by the threads of a CTA with sincos_0pi):
// k = nearest integer to 16 y, d = y - k/16 exactly (both have few enough bits), then
// cos(x_k + d) = C_k cos d - S_k sin d, sin(x_k + d) = S_k cos d + C_k sin d with short Taylor sums.
constexpr int SINCOS_TAB_N = 54;
struct SinCosEntry { double c, s; };
GAPCU_HD void sincos_tab(double y, const SinCosEntry *T, double *sn, double *cs) {
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52
    const double kd = fma(y, 16.0, MAGIC);
    long long kbits;
    memcpy(&kbits, &kd, sizeof kbits);
    const int k = (int)(unsigned int)kbits;
    const double d = fma(kd - MAGIC, -0.0625, y);
    const SinCosEntry e = T[k];
    const double z = d * d;
    double cm1 = fma(z, KC(31), KC(30));
    cm1 = fma(cm1, z, KC(29));
    cm1 = fma(cm1, z, -0.5);
    cm1 *= z;                                             // cos d - 1
    double sp = fma(z, KC(34), KC(33));
    sp = fma(sp, z, KC(32));
    const double sd = fma(sp, d * z, d);                  // sin d
    *cs = fma(-e.s, sd, fma(e.c, cm1, e.c));
    *sn = fma(e.c, sd, fma(e.s, cm1, e.s));
}
static inline void fill_sincos_table(SinCosEntry *T) {
    for (int k = 0; k < SINCOS_TAB_N; k++) { T[k].c = cos(k / 16.0); T[k].s = sin(k / 16.0); }
}


#if defined(__CUDACC__)
// Device variants of exp_neg / sincos_tab whose tables are given as plain 32-bit SHARED addresses
// (__cvta_generic_to_shared): the generic-pointer forms re-derive the shared window base at every call site.
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));   // volatile: stays behind the barrier that follows the table fill
    return v;
}
__device__ __forceinline__ double exp_neg_s(double x, unsigned t32_sa) {
    const double MAGIC = 6755399441055744.0;
    unsigned int xh = (unsigned int)__double2hiint(x);
    xh = xh < 0xc16312d0u ? xh : 0xc16312d0u;
    x = __hiloint2double((int)xh, __double2loint(x));
    const double kd = fma(x, KC(0), MAGIC);
    const double kf = kd - MAGIC;
    const int ki = __double2loint(kd);
    double r = fma(kf, KC(1), x);
    r = fma(kf, KC(2), r);
    double p = fma(r, KC(3), KC(4));
    p = fma(p, r, KC(5));
    p = fma(p, r, KC(6));
    p = fma(p, r, 0.5);
    p = fma(p, r * r, r);
    const double t = lds_f64(t32_sa + ((unsigned)(ki & 31) << 3));
    const double y = fma(t, p, t);
    int m = ki >> 5;
    m = m > -1021 ? m : -1021;
    return __hiloint2double(__double2hiint(y) + (m << 20), __double2loint(y));
}
__device__ __forceinline__ void sincos_tab_s(double y, unsigned trig_sa, double *sn, double *cs) {
    const double MAGIC = 6755399441055744.0;
    const double kd = fma(y, 16.0, MAGIC);
    const int k = __double2loint(kd);
    const double d = fma(kd - MAGIC, -0.0625, y);
    double ec, es;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ec), "=d"(es) : "r"(trig_sa + ((unsigned)k << 4)));
    const double z = d * d;
    double cm1 = fma(z, KC(31), KC(30));
    cm1 = fma(cm1, z, KC(29));
    cm1 = fma(cm1, z, -0.5);
    cm1 *= z;
    double sp = fma(z, KC(34), KC(33));
    sp = fma(sp, z, KC(32));
    const double sd = fma(sp, d * z, d);
    *cs = fma(-es, sd, fma(ec, cm1, ec));
    *sn = fma(ec, sd, fma(es, cm1, es));
}
#endif

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
