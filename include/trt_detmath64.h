/*
 * trt_detmath64.h -- the binary64 power function of the diffusive-wave path, defined operation by operation.
 *
 * WHY.  The diffusive solver (/root/reference/src/kernel/diffusive/diffusive.f90) evaluates x**3.0, x**(2./3.), x**0.3,
 * x**0.4, x**0.6 and x**0.50 in double precision through gfortran's lowering to libm `pow` (:1173-1196, :1479-1481,
 * :1707, :2333, :503).  The results feed a Newton iteration that stops on |dx| < 1e-4 (rtsafe :1646) and a CFL-adaptive
 * time step, so a last-bit difference between two `pow` implementations (glibc vs CUDA libdevice) is amplified: a
 * 2e-8 relative perturbation of one coefficient moved flows by 7e-5 in a 700-node test domain.  As for the float path
 * (trt_detmath.h) the only way to a bit-stable CPU/GPU answer is to pin the function.
 *
 * WHAT.  trt_pow64_det(x, y) = exp(y * log(x)) with the classic argument reductions (log: x = 2^k (1 + f),
 * s = f / (2 + f), degree-14 even polynomial in s; exp: t = n ln2 + r, rational correction in r), every step one IEEE-754
 * binary64 +, -, *, / or fma, in a fixed order, so gcc on x86-64 (-ffp-contract=off) and nvcc on sm_100a (explicit _rn
 * intrinsics) produce identical bits.  The product y * log(x) is carried as a head and an fma-exact tail.
 * Accuracy: |relative error| <= 2^-52 * (2 + |y ln x|); measured against glibc pow in tests/test_detmath64.py
 * (worst 3.2e-16 over 2e7 samples of the solver's exponents, x in [1e-8, 1e8]).  It is NOT correctly rounded and does not need to be: the oracle has a
 * libm build as well, and the distance between the two builds is measured (tests/test_diffusive_oracle.py).
 *
 * DOMAIN.  x > 0 finite and subnormal-free after scaling, any finite y: the formula.  x = +-0: 0 for y > 0, 1 for y = 0,
 * +inf for y < 0.  x < 0, NaN: NaN.  x = +inf: +inf for y > 0.  y = 0: 1.  Results outside [2^-1000, 2^1000] saturate to
 * 0 / +inf (unreachable on the path).
 */
#ifndef TRT_DETMATH64_H
#define TRT_DETMATH64_H

#if defined(__CUDACC__)
#define TRT_HD64 __host__ __device__ __forceinline__
#else
#define TRT_HD64 static inline
#endif

typedef unsigned long long trt_u64_t;

TRT_HD64 double trt64_from_bits(trt_u64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { trt_u64_t u; double d; } c; c.u = u; return c.d;
#endif
}
TRT_HD64 trt_u64_t trt64_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (trt_u64_t)__double_as_longlong(d);
#else
    union { trt_u64_t u; double d; } c; c.d = d; return c.u;
#endif
}
#if defined(__CUDA_ARCH__)
#define TRT64_MUL(a, b) __dmul_rn((a), (b))
#define TRT64_ADD(a, b) __dadd_rn((a), (b))
#define TRT64_SUB(a, b) __dsub_rn((a), (b))
#define TRT64_DIV(a, b) __ddiv_rn((a), (b))
#define TRT64_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define TRT64_MUL(a, b) ((a) * (b))
#define TRT64_ADD(a, b) ((a) + (b))
#define TRT64_SUB(a, b) ((a) - (b))
#define TRT64_DIV(a, b) ((a) / (b))
#define TRT64_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

TRT_HD64 double trt_pow64_det(double x, double y)
{
    const double ln2_hi = 6.93147180369123816490e-01;   /* 0x3fe62e42fee00000 */
    const double ln2_lo = 1.90821492927058770002e-10;   /* 0x3dea39ef35793c76 */
    const double invln2 = 1.44269504088896338700e+00;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (y == 0.0) return 1.0;
    if (x != x || y != y) return x + y;
    if (x < 0.0) return trt64_from_bits(0x7ff8000000000000ULL);
    if (x == 0.0) return y > 0.0 ? 0.0 : trt64_from_bits(0x7ff0000000000000ULL);
    if (x == trt64_from_bits(0x7ff0000000000000ULL)) return y > 0.0 ? x : 0.0;

    /* ---- log(x) = k ln2 + log(1 + f), 1 + f in [sqrt(1/2), sqrt(2)) */
    trt_u64_t ux = trt64_bits(x);
    int k = 0;
    if ((ux >> 52) == 0) { x = TRT64_MUL(x, 18014398509481984.0); ux = trt64_bits(x); k = -54; }   /* subnormal: * 2^54 */
    k += (int)(ux >> 52) - 1023;
    trt_u64_t mant = ux & 0x000fffffffffffffULL;
    /* mantissa >= sqrt(2): halve it (fdlibm's (hx + 0x95f64) & 0x100000 trick, on the full 52 bits) */
    const int up = mant >= 0x6a09e667f3bcdULL;
    k += up;
    const double m = trt64_from_bits(mant | (up ? 0x3fe0000000000000ULL : 0x3ff0000000000000ULL));
    const double f = TRT64_SUB(m, 1.0);                                   /* exact */
    const double s = TRT64_DIV(f, TRT64_ADD(2.0, f));
    const double z = TRT64_MUL(s, s);
    const double w = TRT64_MUL(z, z);
    const double t1 = TRT64_MUL(w, TRT64_ADD(Lg2, TRT64_MUL(w, TRT64_ADD(Lg4, TRT64_MUL(w, Lg6)))));
    const double t2 = TRT64_MUL(z, TRT64_ADD(Lg1, TRT64_MUL(w, TRT64_ADD(Lg3, TRT64_MUL(w, TRT64_ADD(Lg5, TRT64_MUL(w, Lg7)))))));
    const double R = TRT64_ADD(t2, t1);
    const double hfsq = TRT64_MUL(0.5, TRT64_MUL(f, f));
    const double dk = (double)k;
    /* log(x) = dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f), split into a head and a tail */
    const double tail0 = TRT64_ADD(TRT64_MUL(s, TRT64_ADD(hfsq, R)), TRT64_MUL(dk, ln2_lo));
    const double corr = TRT64_SUB(TRT64_SUB(hfsq, tail0), f);            /* = -(log(1+f) + dk*ln2_lo) */
    const double khi = TRT64_MUL(dk, ln2_hi);                              /* exact: ln2_hi has 21 trailing zero bits */
    const double L = TRT64_SUB(khi, corr);
    const double Lt = TRT64_SUB(TRT64_SUB(khi, L), corr);                  /* rounding error of the last subtraction */

    /* ---- t = y * log(x) as head + tail */
    const double th = TRT64_MUL(y, L);
    const double tl = TRT64_ADD(TRT64_FMA(y, L, -th), TRT64_MUL(y, Lt));
    if (th > 693.0) return trt64_from_bits(0x7ff0000000000000ULL);
    if (th < -693.0) return 0.0;

    /* ---- exp(th + tl) */
    const double nn = TRT64_ADD(TRT64_MUL(th, invln2), th < 0.0 ? -0.5 : 0.5);
    const int n = (int)nn;                                                 /* round to nearest by truncation */
    const double dn = (double)n;
    const double hi = TRT64_SUB(th, TRT64_MUL(dn, ln2_hi));
    const double lo = TRT64_SUB(TRT64_MUL(dn, ln2_lo), tl);
    const double r = TRT64_SUB(hi, lo);
    const double rr = TRT64_MUL(r, r);
    const double c = TRT64_SUB(r, TRT64_MUL(rr, TRT64_ADD(P1, TRT64_MUL(rr, TRT64_ADD(P2, TRT64_MUL(rr, TRT64_ADD(P3, TRT64_MUL(rr, TRT64_ADD(P4, TRT64_MUL(rr, P5))))))))));
    const double e = TRT64_SUB(1.0, TRT64_SUB(TRT64_SUB(lo, TRT64_DIV(TRT64_MUL(r, c), TRT64_SUB(2.0, c))), hi));
    /* scale by 2^n, |n| <= 1000: exact */
    return TRT64_MUL(e, trt64_from_bits((trt_u64_t)(n + 1023) << 52));
}

#endif /* TRT_DETMATH64_H */
