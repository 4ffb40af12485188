/*
 * trt_detmath.h -- the numerics contract of the B200 routing path: one powf, defined bit by bit.
 *
 * WHY.  The reference evaluates x**(2/3), x**(5/3), x**0.5 and x**1.5 in real(4)
 * (src/kernel/muskingum/MCsingleSegStime_f2py_NOLOOP.f90:168-169, :251-257, :261-264, :328-329,
 * :356-363; src/kernel/reservoir/Level_Pool/module_levelpool.F:303,307,324,347,371) through
 * gfortran's lowering to libm `powf`.  The result feeds the termination test of the secant
 * iteration (:83), so a 1-ulp difference between two powf implementations can change the
 * iteration count of a lane and move its outflow by ~1e-3 relative -- far outside the 1e-5
 * parity bar.  libm powf is faithful but not correctly rounded and differs between glibc
 * versions / FMA ifunc variants, and CUDA's powf is different again.  The only way to get a
 * per-segment, per-step bit-stable answer on CPU and GPU is to pin the function itself.
 *
 * WHAT.  trt_powf_det(x, y) = (float) 2^( y * log2(x) ) with log2 and 2^ evaluated in IEEE
 * binary64 using only +, *, fma and integer bit moves (all exactly specified by IEEE-754), so
 * gcc on x86-64 and nvcc on sm_100a produce identical bits.  Accuracy of the binary64
 * intermediate is ~2^-49 relative for the argument range of the path, i.e. the result is the
 * CORRECTLY ROUNDED powf except on ~1e-7 of inputs (tests/test_detmath.py measures this against
 * mpmath) -- tighter than any libm the reference could have been linked against (glibc 2.28+:
 * 0.82 ulp, e_powf.c).
 *
 *   log2(x): x = 2^k * z, z in [0.6875, 1.375); 128-bin table {invc, logc = -log2(invc)} with
 *            invc <= 28 significant bits so r = z*invc - 1 is exact; log2(1+r) by a degree-7
 *            Taylor polynomial (|r| <= 2^-7); the bin holding 1.0 has invc = 1, logc = 0.
 *   2^t    : t = n/32 + g, |g| <= 1/64; 2^(j/32) table (j = n mod 32), degree-6 Taylor
 *            polynomial for 2^g - 1, exponent n div 32 applied by an exact power-of-two scale;
 *            one final binary64 -> binary32 rounding (handles overflow / subnormals).
 *
 * DOMAIN.  All float x; finite y > 0 that is NOT an integer (the path uses 2/3, 5/3, 1/2, 3/2).
 *          x < 0 -> NaN, x = +-0 -> +0, x = +inf -> +inf, NaN -> NaN  (powf semantics for such y,
 *          except powf(-inf, y) = +inf which this returns as NaN; unreachable on the path).
 *
 * The tables are passed by pointer so that device code can stage them in shared memory
 * (device: tl must be 16-byte aligned, each {invc, logc} pair is fetched with one 128-bit load).
 * Included by the CUDA kernels (t-route_b200/csrc) and by the CPU oracle (oracle/), which also
 * has a libm-powf build used to measure how far the platform libm is from this definition.
 */
#ifndef TRT_DETMATH_H
#define TRT_DETMATH_H

#include "trt_detmath_tables.h"

#if defined(__CUDACC__)
#define TRT_HD __host__ __device__ __forceinline__
#else
#define TRT_HD static inline
#endif

typedef unsigned long long trt_u64;

TRT_HD double trt_u2d(trt_u64 u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { trt_u64 u; double d; } c; c.u = u; return c.d;
#endif
}
TRT_HD trt_u64 trt_d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (trt_u64)__double_as_longlong(d);
#else
    union { trt_u64 u; double d; } c; c.d = d; return c.u;
#endif
}
/* the three binary64 operations of the contract; never contracted, never reassociated */
TRT_HD double trt_dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
TRT_HD double trt_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
TRT_HD double trt_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

/* tl: TRT_LOG2_TAB_N {invc, logc} pairs; te: TRT_EXP2_TAB_N entries; both as binary64 bit patterns */

/* Polynomial coefficients.  Device code may define TRT_DEVICE_COEF as the name of a `__constant__ double[13]` holding
 * {A1..A7, B1..B6}: a DFMA then takes the coefficient straight from the constant bank instead of building the 64-bit
 * immediate with two moves per use (174 of the ~2050 instructions per segment-timestep were such moves). */
#if defined(__CUDA_ARCH__) && defined(TRT_DEVICE_COEF)
#define TRT_LOG2_A(i) (TRT_DEVICE_COEF[(i) - 1])
#define TRT_EXP2_B(i) (TRT_DEVICE_COEF[6 + (i)])
#else
#define TRT_LOG2_A(i) trt_u2d(TRT_LOG2_A##i##_BITS)
#define TRT_EXP2_B(i) trt_u2d(TRT_EXP2_B##i##_BITS)
#endif

/* log2(x) in binary64 for finite x > 0 */
TRT_HD double trt_log2_pos(float x, const trt_u64* tl) {
    const trt_u64 ix  = trt_d2u((double)x);
    const trt_u64 tmp = ix - 0x3fe6000000000000ULL;                /* OFF = bits(0.6875) */
    const int i       = (int)((tmp >> 45) & (TRT_LOG2_TAB_N - 1));
    const long long k = (long long)tmp >> 52;                      /* arithmetic shift */
    const double z    = trt_u2d(ix - (tmp & 0xfff0000000000000ULL));
#if defined(__CUDA_ARCH__)
    const ulonglong2 tle = *reinterpret_cast<const ulonglong2*>(tl + 2 * i);   /* tl must be 16-byte aligned */
    const double invc = trt_u2d(tle.x);
    const double logc = trt_u2d(tle.y);
#else
    const double invc = trt_u2d(tl[2 * i]);
    const double logc = trt_u2d(tl[2 * i + 1]);
#endif
    const double r    = trt_dfma(z, invc, -1.0);                   /* exact */
    double p = TRT_LOG2_A(7);
    p = trt_dfma(p, r, TRT_LOG2_A(6));
    p = trt_dfma(p, r, TRT_LOG2_A(5));
    p = trt_dfma(p, r, TRT_LOG2_A(4));
    p = trt_dfma(p, r, TRT_LOG2_A(3));
    p = trt_dfma(p, r, TRT_LOG2_A(2));
    p = trt_dfma(p, r, TRT_LOG2_A(1));
    const double l     = trt_dadd((double)k, logc);
    return trt_dfma(p, r, l);
}

/* (float) 2^(y * log2x) with one final rounding */
TRT_HD float trt_exp2_scaled(double log2x, float y, const trt_u64* te) {
    double t = trt_dmul((double)y, log2x);
    /* 2^t overflows binary32 for t >= 128 and rounds to +0 for t <= -150.  Clamping t to [-200, 200] keeps the exponent
     * arithmetic below in range and lets the final binary64 -> binary32 conversion produce +inf / +0 by itself: no
     * branches (a branch here is two per power in the device code), same results as returning those values early. */
#ifdef TRT_POW_BRANCHY
    if (t >= 130.0) return trt_u2d(0x7ff0000000000000ULL);
    if (t <= -160.0) return 0.0f;
#else
    t = t > 200.0 ? 200.0 : t;
    t = t < -200.0 ? -200.0 : t;
#endif
    const double SH = 6755399441055744.0;                          /* 1.5 * 2^52 */
    const double u  = trt_dmul(t, 32.0);                           /* exact */
    double kd       = trt_dadd(u, SH);                             /* round to nearest integer */
    const trt_u64 ki = trt_d2u(kd);
    kd              = trt_dadd(kd, -SH);
    const double g  = trt_dmul(trt_dadd(u, -kd), 0.03125);         /* exact, |g| <= 1/64 */
    const int n     = (int)(unsigned int)ki;                       /* low 32 bits: two's complement n */
    const int j     = n & (TRT_EXP2_TAB_N - 1);
    const int q     = n >> 5;                                      /* arithmetic shift */
    double e = TRT_EXP2_B(6);
    e = trt_dfma(e, g, TRT_EXP2_B(5));
    e = trt_dfma(e, g, TRT_EXP2_B(4));
    e = trt_dfma(e, g, TRT_EXP2_B(3));
    e = trt_dfma(e, g, TRT_EXP2_B(2));
    e = trt_dfma(e, g, TRT_EXP2_B(1));
    const double w  = trt_dmul(e, g);                              /* 2^g - 1 */
    const double s  = trt_u2d(te[j]);
    const double m  = trt_dfma(s, w, s);                           /* 2^(j/32 + g) in [1, 2) */
    const double sc = trt_u2d((trt_u64)(long long)(q + 1023) << 52); /* 2^q, q in [-201, 200] */
    return (float)trt_dmul(m, sc);                                 /* single rounding to binary32 */
}

/* Arguments that never reach the logarithm (x <= 0, NaN, +inf) are replaced by 1.0 for the evaluation and their result
 * is selected afterwards: straight-line code, no divergence between the lanes of a warp. */
TRT_HD float trt_pow_special(float x, int* special) {
    *special = 1;
    if (!(x > 0.0f)) {                       /* x <= 0 or NaN */
        if (x == 0.0f) return 0.0f;          /* +-0 ** (y>0) = +0 */
        return x != x ? x : trt_u2d(0x7ff8000000000000ULL) /* NaN */;
    }
    if (x > 3.402823466e+38f) return x;      /* +inf */
    *special = 0;
    return 0.0f;
}

TRT_HD float trt_powf_det(float x, float y, const trt_u64* tl, const trt_u64* te) {
    int special;
    const float r = trt_pow_special(x, &special);
#ifdef TRT_POW_BRANCHY
    if (special) return r;
#endif
    const float v = trt_exp2_scaled(trt_log2_pos(special ? 1.0f : x, tl), y, te);
    return special ? r : v;
}

/* x**y1 and x**y2 from ONE logarithm: bit-identical to two trt_powf_det calls (both are pure functions of the same
 * log2(x)), at 2/3 of the cost.  The Muskingum-Cunge celerity needs R**(2/3) and R**(5/3) of the same R. */
TRT_HD void trt_powf_det2(float x, float y1, float y2, float* o1, float* o2, const trt_u64* tl, const trt_u64* te) {
    int special;
    const float r = trt_pow_special(x, &special);
#ifdef TRT_POW_BRANCHY
    if (special) { *o1 = r; *o2 = r; return; }
#endif
    const double l = trt_log2_pos(special ? 1.0f : x, tl);
    const float v1 = trt_exp2_scaled(l, y1, te);
    const float v2 = trt_exp2_scaled(l, y2, te);
    *o1 = special ? r : v1;
    *o2 = special ? r : v2;
}

/* (float) e^x for a binary64 x, as (float) 2^(x * log2 e) through the same exp2 core: the decay weight of streamflow
 * nudging, exp(|minutes| / -a) (simple_da.pyx:120, a libm double `exp` in the reference).  Within 1 float ulp of a
 * correctly rounded result, identical on CPU and GPU. */
TRT_HD float trt_expf_det(double x, const trt_u64* te) {
    const double log2e = trt_u2d(0x3ff71547652b82feULL);           /* 1.4426950408889634 */
    return trt_exp2_scaled(trt_dmul(x, log2e), 1.0f, te);
}

#endif /* TRT_DETMATH_H */
