// fastmath.cuh -- FP64 exp / sincos specialised to the argument ranges of this path.
//
// The wACSF kernels spend most of their FP64 work in exp(-alpha*s) (s >= 0) and in
// cos/sin(pi~ r/Rc) with r <= Rc.  The CUDA library versions carry range checks,
// huge-argument reduction and denormal handling that are dead weight here (ncu:
// ~50 SASS instructions per exp, ~70 per sincos).  These versions keep full double
// accuracy (<= ~2 ulp; tests/test_fastmath.py checks them against libm on the host
// with the same source) on the restricted domains:
//   exp_neg(x)      x <= 0  (clamped at -700: returns ~1e-304 rather than denormals)
//   sincos_0pi(y)   0 <= y <= 3.3
#pragma once
#include <math.h>

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

// exp(x) for x <= 0.  x = (32 m + j) ln2/32 + r with |r| <= ln2/64, so
// exp(x) = 2^m * T[j] * (1 + r + r^2/2! + ... + r^6/6!)        (next term < 4e-18)
GAPCU_HD double exp_neg(double x, const double *T32) {
    x = fmax(x, -700.0);
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: rounds to nearest integer
    const double kd = fma(x, 46.16624130844683, MAGIC);  // 32 / ln2
    const double kf = kd - MAGIC;
    long long kbits;
    memcpy(&kbits, &kd, sizeof kbits);
    const int ki = (int)(unsigned int)kbits;             // low word of kd holds the integer (two's complement)
    double r = fma(kf, -0.021660849392446835, x);        // ln2/32, high 36 bits: k*hi is exact
    r = fma(kf, -5.145609244655338e-14, r);              // ln2/32, low part
    double p = fma(r, 0.001388888888888889, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r * r, r);                                // expm1(r)
    const double t = T32[ki & 31];
    double y = fma(t, p, t);
    long long yb;
    memcpy(&yb, &y, sizeof yb);
    yb += (long long)(ki >> 5) << 52;                    // * 2^m (y in [1,2.1), m >= -1010: stays normal)
    memcpy(&y, &yb, sizeof y);
    return y;
}

// sin(y), cos(y) for 0 <= y <= 3.3: quadrant q in {0,1,2}, t = y - q*pi/2 in [-pi/4, pi/4]
GAPCU_HD void sincos_0pi(double y, double *sn, double *cs) {
    const double q = rint(y * 0.6366197723675814);       // 2/pi
    double t = fma(q, -1.5707963267948912, y);           // pi/2 high part (q*hi exact)
    t = fma(q, -5.390302858158119e-15, t);
    const double z = t * t;
    double s = fma(z, 2.8114572543455206e-15, -7.647163731819816e-13);
    s = fma(s, z, 1.6059043836821613e-10);
    s = fma(s, z, -2.505210838544172e-08);
    s = fma(s, z, 2.7557319223985893e-06);
    s = fma(s, z, -0.0001984126984126984);
    s = fma(s, z, 0.008333333333333333);
    s = fma(s, z, -0.16666666666666666);
    s = fma(s * z, t, t);                                // sin(t)
    double c = fma(z, 4.779477332387385e-14, -1.1470745597729725e-11);
    c = fma(c, z, 2.08767569878681e-09);
    c = fma(c, z, -2.755731922398589e-07);
    c = fma(c, z, 2.48015873015873e-05);
    c = fma(c, z, -0.001388888888888889);
    c = fma(c, z, 0.041666666666666664);
    c = fma(c, z, -0.5);
    c = fma(c, z, 1.0);                                  // cos(t)
    // q = 0: (s, c); q = 1: (c, -s); q = 2: (-s, -c)
    const bool odd = (q == 1.0);
    const double sgn = (q == 2.0) ? -1.0 : 1.0;
    *sn = odd ? c : sgn * s;
    *cs = odd ? -s : sgn * c;
}

}  // namespace gapcu
