/*
 * cvmath.h -- scalar FP64 building blocks of the likelihood kernels.
 *
 * Everything here is `__host__ __device__` so that the same arithmetic that runs in the sm_100a
 * kernels can be unit-tested on the build container's CPU (tests/host_math, never part of the
 * product path).  All operations whose rounding matters are written with explicit
 * round-to-nearest primitives (cv_mul/cv_add/cv_fma), which nvcc never contracts or reorders.
 *
 * Reference lines are relative to /root/reference.
 */
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CV_HD __host__ __device__ __forceinline__
#else
#define CV_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define cv_mul(a, b) __dmul_rn((a), (b))
#define cv_add(a, b) __dadd_rn((a), (b))
#define cv_sub(a, b) __dsub_rn((a), (b))
#define cv_fma(a, b, c) __fma_rn((a), (b), (c))
#define cv_div(a, b) __ddiv_rn((a), (b))
#else
/* host builds are compiled with -ffp-contract=off */
#define cv_mul(a, b) ((a) * (b))
#define cv_add(a, b) ((a) + (b))
#define cv_sub(a, b) ((a) - (b))
#define cv_fma(a, b, c) fma((a), (b), (c))
#define cv_div(a, b) ((a) / (b))
#endif

CV_HD uint64_t cv_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    union {
        double d;
        uint64_t u;
    } c;
    c.d = x;
    return c.u;
#endif
}

CV_HD double cv_from_bits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union {
        double d;
        uint64_t u;
    } c;
    c.u = u;
    return c.d;
#endif
}

/* ----------------------------------------------------------------------------------------- */
/* double-double: value = hi + lo, |lo| <= ulp(hi)/2                                           */
/* ----------------------------------------------------------------------------------------- */
struct cv_dd {
    double hi, lo;
};

CV_HD cv_dd cv_two_sum(double a, double b)
{
    double s = cv_add(a, b);
    double bb = cv_sub(s, a);
    double err = cv_add(cv_sub(a, cv_sub(s, bb)), cv_sub(b, bb));
    cv_dd r = {s, err};
    return r;
}

/* requires |a| >= |b| (or a == 0) */
CV_HD cv_dd cv_fast_two_sum(double a, double b)
{
    double s = cv_add(a, b);
    double err = cv_sub(b, cv_sub(s, a));
    cv_dd r = {s, err};
    return r;
}

CV_HD cv_dd cv_two_prod(double a, double b)
{
    double p = cv_mul(a, b);
    double err = cv_fma(a, b, -p);
    cv_dd r = {p, err};
    return r;
}

CV_HD cv_dd cv_dd_mul(cv_dd a, cv_dd b)
{
    cv_dd p = cv_two_prod(a.hi, b.hi);
    double cross = cv_fma(a.hi, b.lo, cv_mul(a.lo, b.hi));
    return cv_fast_two_sum(p.hi, cv_add(p.lo, cross));
}

CV_HD cv_dd cv_dd_mul_d(cv_dd a, double b)
{
    cv_dd p = cv_two_prod(a.hi, b);
    return cv_fast_two_sum(p.hi, cv_fma(a.lo, b, p.lo));
}

CV_HD cv_dd cv_dd_add(cv_dd a, cv_dd b)
{
    cv_dd s = cv_two_sum(a.hi, b.hi);
    return cv_fast_two_sum(s.hi, cv_add(s.lo, cv_add(a.lo, b.lo)));
}

CV_HD cv_dd cv_dd_add_d(cv_dd a, double b)
{
    cv_dd s = cv_two_sum(a.hi, b);
    return cv_fast_two_sum(s.hi, cv_add(s.lo, a.lo));
}

/* x**n for a small non-negative integer n, rounded once at the end: the value libm's pow()
 * returns whenever pow() is correctly rounded (models.py:76-78 `(1.0 - err) ** (k - s)`,
 * `err ** s`).  pow(x, 0) = 1 for every x, as in C and Python. */
CV_HD double cv_pow_uint(double x, int n)
{
    if (n == 0)
        return 1.0;
    if (x != x)
        return x;
    cv_dd base = {x, 0.0};
    cv_dd acc = {1.0, 0.0};
    bool started = false;
    while (n > 0) {
        if (n & 1) {
            acc = started ? cv_dd_mul(acc, base) : base;
            started = true;
        }
        n >>= 1;
        if (n > 0)
            base = cv_dd_mul(base, base);
    }
    double r = cv_add(acc.hi, acc.lo);
    /* the error-free transforms above break down when intermediate error terms underflow;
     * results that small only ever feed weights that evaluate to exactly zero */
    return r;
}

/* ----------------------------------------------------------------------------------------- */
/* exp() rounded the way the libm that ran the reference rounds it (glibc >= 2.28 on an FMA     */
/* capable x86-64: the table-driven algorithm with N = 128, contracted with FMAs).  Valid for   */
/* -500 < x <= 0, the only arguments `1.0 - exp(-l)` (models.py:87, :221) needs; verified bit   */
/* for bit against the container's libm in tests/test_math_host.py.                            */
/* ----------------------------------------------------------------------------------------- */
#include "exp_table.h"
static const unsigned long long cv_exp_table_host[256] = {CV_EXP_TABLE_VALUES};
#ifdef __CUDACC__
static __device__ const unsigned long long cv_exp_table_dev[256] = {CV_EXP_TABLE_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define CV_EXP_TABLE cv_exp_table_dev
#else
#define CV_EXP_TABLE cv_exp_table_host
#endif

CV_HD double cv_exp_libm(double x)
{
    const double inv_ln2_n = 0x1.71547652b82fep7;
    const double shift = 0x1.8p52;
    const double neg_ln2_hi_n = -0x1.62e42fefa0000p-8;
    const double neg_ln2_lo_n = -0x1.cf79abc9e3b3ap-47;
    const double c2 = 0x1.ffffffffffdbdp-2, c3 = 0x1.555555555543cp-3;
    const double c4 = 0x1.55555cf172b91p-5, c5 = 0x1.1111167a4d017p-7;
    if (fabs(x) < 0x1p-54)
        return cv_add(1.0, x);
    double z = cv_mul(inv_ln2_n, x);
    double kd = cv_add(z, shift);
    uint64_t ki = cv_bits(kd);
    kd = cv_sub(kd, shift);
    double r = cv_fma(kd, neg_ln2_lo_n, cv_fma(kd, neg_ln2_hi_n, x));
    uint64_t idx = 2 * (ki % 128);
    uint64_t top = ki << 45;
    double tail = cv_from_bits(CV_EXP_TABLE[idx]);
    uint64_t sbits = CV_EXP_TABLE[idx + 1] + top;
    double r2 = cv_mul(r, r);
    double lo_poly = cv_fma(r2, cv_fma(r, c3, c2), cv_add(tail, r));
    double tmp = cv_fma(cv_mul(r2, r2), cv_fma(r, c5, c4), lo_poly);
    double scale = cv_from_bits(sbits);
    return cv_fma(scale, tmp, scale);
}

/* `1.0 - exp(-rate)` of models.py:87 / :221.  Above 38 the exponential is below 2^-54 and
 * the difference is exactly 1. */
CV_HD double cv_one_minus_exp_neg(double rate)
{
    if (rate != rate)
        return rate;
    if (rate >= 38.0)
        return 1.0;
    if (rate <= -500.0)
        return -INFINITY;
    return cv_sub(1.0, cv_exp_libm(-rate));
}

/* ----------------------------------------------------------------------------------------- */
/* log(x) as a double-double for a positive normal x: absolute error ~1e-18 * |log x|.          */
/* Multiplied by a bin index of a few thousand it still leaves the exponent of the Poisson      */
/* term good to ~1e-15.  log(m) = 2 atanh(z), z = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)).       */
/* ----------------------------------------------------------------------------------------- */
CV_HD cv_dd cv_log_dd(double x)
{
    const double ln2_hi = 0x1.62e42fefa3000p-1;    /* 41 significant bits: e * ln2_hi is exact */
    const double ln2_lo = 0x1.3de6af278ece6p-42;   /* ln2 - ln2_hi */
    uint64_t u = cv_bits(x);
    int e = (int)((u >> 52) & 0x7ff) - 1023;
    double m = cv_from_bits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL); /* [1,2) */
    if (m > 1.4142135623730951) {
        m = cv_mul(m, 0.5);
        e += 1;
    }
    double f = cv_sub(m, 1.0);              /* exact */
    cv_dd g = cv_two_sum(m, 1.0);           /* m + 1 */
    double zh = cv_div(f, g.hi);
    double rem = cv_fma(-zh, g.hi, f);      /* exact remainder f - zh*g.hi */
    rem = cv_fma(-zh, g.lo, rem);
    double zl = cv_div(rem, g.hi);
    double w = cv_mul(zh, zh);
    /* sum_{n>=1} w^n / (2n+1), n = 1..13 */
    double p = 1.0 / 27.0;
    p = cv_fma(p, w, 1.0 / 25.0);
    p = cv_fma(p, w, 1.0 / 23.0);
    p = cv_fma(p, w, 1.0 / 21.0);
    p = cv_fma(p, w, 1.0 / 19.0);
    p = cv_fma(p, w, 1.0 / 17.0);
    p = cv_fma(p, w, 1.0 / 15.0);
    p = cv_fma(p, w, 1.0 / 13.0);
    p = cv_fma(p, w, 1.0 / 11.0);
    p = cv_fma(p, w, 1.0 / 9.0);
    p = cv_fma(p, w, 1.0 / 7.0);
    p = cv_fma(p, w, 1.0 / 5.0);
    p = cv_fma(p, w, 1.0 / 3.0);
    p = cv_mul(p, w);
    /* 2z(1 + p) with z = zh + zl; the zl*p cross term is below 1e-33 */
    double corr = cv_fma(zh, p, zl);
    cv_dd lm = cv_fast_two_sum(cv_mul(2.0, zh), cv_mul(2.0, corr));
    double ed = (double)e;
    cv_dd r = cv_two_sum(cv_mul(ed, ln2_hi), lm.hi);
    return cv_fast_two_sum(r.hi, cv_add(r.lo, cv_fma(ed, ln2_lo, lm.lo)));
}

/* ----------------------------------------------------------------------------------------- */
/* log(x) for the GEMM epilogue (safe_log of a bin probability, utils.py:32-35): table driven, */
/* 128 intervals over [0.6875, 1.375), x = 2^k z, r = z * invc - 1 with |r| < 2^-7.6, then     */
/* log x = k ln2 + logc + log1p(r), log1p as a degree-7 polynomial.  Absolute error below      */
/* 2e-16 (1 + |log x|): far inside what the count-weighted sum needs.  `tab` holds 128 pairs   */
/* (invc, logc) built by cv_log_table.  Zero, negative, subnormal and non-finite arguments     */
/* take libm's log (0 -> -inf, the caller maps p <= 0 to -inf first).                          */
/* ----------------------------------------------------------------------------------------- */
#define CV_LOG_N 128
#define CV_LOG_OFF 0x3fe6000000000000ULL

/* kshift: log(x * 2^-kshift) -- exact, the shift joins the exponent */
CV_HD double cv_log_tab(double x, const double *tab, int kshift = 0)
{
    if (!(x >= 0x1p-1022) || !(x < INFINITY))
        return cv_sub(log(x), cv_mul((double)kshift, 0x1.62e42fefa39efp-1));
    const uint64_t ix = cv_bits(x);
    const uint64_t tmp = ix - CV_LOG_OFF;
    const int i = (int)((tmp >> 45) & (CV_LOG_N - 1));
    const int k = (int)((int64_t)tmp >> 52) - kshift;
    const double z = cv_from_bits(ix - (tmp & 0xfff0000000000000ULL));
    const double invc = tab[2 * i], logc = tab[2 * i + 1];
    const double r = cv_fma(z, invc, -1.0);
    const double kd = (double)k;
    const double hi = cv_fma(kd, 0x1.62e42fefa3800p-1, logc); /* ln2 to 42 bits: kd * ln2hi exact */
    double p = cv_fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = cv_fma(r, p, 1.0 / 5.0);
    p = cv_fma(r, p, -1.0 / 4.0);
    p = cv_fma(r, p, 1.0 / 3.0);
    p = cv_fma(r, p, -0.5);
    const double r2 = cv_mul(r, r);
    const double lo = cv_fma(r2, p, cv_mul(kd, 0x1.ef35793c76730p-45)); /* ln2 - ln2hi */
    return cv_add(hi, cv_add(r, lo));
}

/* host: the 128 (invc, logc) pairs, logc = -log(invc) of the ROUNDED invc */
static inline void cv_log_table(double *tab)
{
    for (int i = 0; i < CV_LOG_N; i++) {
        long double u = ((long double)i + 0.5L) / CV_LOG_N;
        long double c = u < 0.625L ? 0.5L * (1.375L + u) : 0.375L + u;
        double invc = (double)(1.0L / c);
        tab[2 * i] = invc;
        tab[2 * i + 1] = (double)(-logl((long double)invc));
    }
}
