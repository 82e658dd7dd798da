/*
 * cvmodel.h -- per-point / per-term arithmetic of the CovEst mixture model, shared by the
 * sm_100a kernels (loglik_kernel.cu) and the test-only host emulation (tests/host_math).
 *
 * Formulation.  For one parameter point the reference evaluates
 *
 *     p_j = sum_o b(o) * sum_s a_os * tp(o*l_s, j)                    (models.py:235-241)
 *     tp(L, j) = (prod_{i<=j} L/i) / D(L)                               (covest_poissonmodule.c:7-35)
 *
 * Every (o, s) pair is one *term* with rate L = o*l_s and weight w = b(o)*a_os.  Its value at
 * bin j is w * L^j / j! / D(L) = exp(j*log L - lgamma(j+1) - lin) * w / Dred.  The kernels
 * evaluate that exponential once per (term, chain of bins) -- the *seed*, in double-double so
 * the large cancelling parts j*log L, lgamma(j+1) and L keep ~1e-15 absolute accuracy -- and
 * walk the following bins with the recurrence value(j+1) = value(j) * L / (j+1), which is the
 * same product the reference's inner loop forms.
 *
 * D(L) is the reference's denominator *as implemented* (c:25-31), not e^L - 1:
 *     n = number of times 200 can be taken off L while it stays > 200,  r = L - 200 n
 *     D = e^(200 n) * (e^r - 1)     if r > 1e-8
 *     D = e^(200 n) * L             otherwise (the ORIGINAL L, c:19 `p3 = l`)
 */
#pragma once
#include "cvmath.h"

/* exp(CV_SCALE_LOG) multiplies every seed so that terms that are still far below the double
 * range at the head of a chain, but grow inside it, keep full precision; the epilogue divides it
 * out again.  420 leaves room both ways for chains of up to 64 bins and rates up to the
 * reference's own overflow limit (~11 360). */
#define CV_SCALE_LOG 420.0

/* Probabilities leave the accumulators scaled by 2^128 (slot_mult carries the factor): they stay
 * normal numbers down to p = 2^-1150, so a kernel can tell three classes of a bin with a count:
 *
 *   p >= 2^-1030   the reference rounds every term to a double (c:33) and every product
 *                  (models.py:236-239); where those are subnormal each rounding moves p by up to
 *                  2^-1075, 3 T of them (T terms) by 3 T 2^-1075 / p relatively, and the
 *                  log-likelihood -- whose magnitude is at least h |log p| > 693 h for this very bin
 *                  -- by less than 3 T 2^-45 / 693 = 1e-12 relatively for T up to 8192 terms: the
 *                  fast arithmetic of the kernels stands
 *   p <  2^-1080   the reference's value is exactly 0 (a non-zero result needs a product that
 *                  rounds to at least one unit 2^-1074 of the subnormal grid, which takes a true
 *                  value of at least 2^-1077): log-likelihood -inf
 *   in between     the reference's per-term roundings to the subnormal grid decide the value
 *                  (one unit is ln 2 in log p): the bin contributes the sentinel CV_BAND_LOG, which
 *                  makes the point's value < CV_BAND_LL, and such points are re-evaluated term by
 *                  term, rounding where the reference rounds (faithful.cu).  They are far-off
 *                  points (a bin with a count is 700 nats off), about one in a hundred of a wide
 *                  candidate box -- the ones whose last counted bins run out of the double range. */
#define CV_PSCALE 0x1p128
#define CV_PUNSCALE 0x1p-128
#define CV_PSCALE_EXP 128
#define CV_PSCALE_LOG 0x1.62e42fefa39efp+6 /* 128 ln 2 */
#define CV_P_ZERO 0x1p-952  /* scaled: p < 2^-1080 */
#define CV_P_BAND 0x1p-902  /* scaled: p < 2^-1030 */
#define CV_BAND_LOG (-1.0e280)
#define CV_BAND_LL (-1.0e270)
#define CV_DEAD_TERM (-1.0e30)
#define CV_MAX_ERR 64

struct CvTerm {
    double lam;    /* rate o*l_s (1 for a term that contributes nothing) */
    double lh, ll; /* log(lam) as double-double */
    double lin;    /* the part of log D(lam) that is linear: lam, or 200 n (exact in double) */
    double f;      /* w / Dred with D(lam) = e^lin * Dred; 0 for a dead term, NaN propagates */
};

/* models.py:60-69 fit_to_bounds for one coordinate; NaN bound = open */
CV_HD double cv_clip(double v, double lo, double hi)
{
    if (lo == lo && v < lo)
        return lo;
    if (hi == hi && v > hi)
        return hi;
    return v;
}

/* models.py:71-72 */
CV_HD double cv_kmer_coverage(double c, int k, int r)
{
    return cv_div(cv_mul(c, (double)(r - k + 1)), (double)r);
}

/* models.py:75-79: l_s = c_k * 3**-s * (1-err)**(k-s) * err**s, multiplied left to right.
 * pow3_neg_s is the host libm's pow(3.0, -s) (1.0 for s = 0). */
CV_HD double cv_error_class_rate(double ck, double pow3_neg_s, double err, int k, int s)
{
    double keep = cv_pow_uint(cv_sub(1.0, err), k - s);
    double miss = cv_pow_uint(err, s);
    return cv_mul(cv_mul(cv_mul(ck, pow3_neg_s), keep), miss);
}

/* models.py:193-208 get_b_o.  two = (1-q1)*q2, many = (1-q1)*(1-q2)*q, base = 1-q. */
CV_HD double cv_copy_weight(int o, double q1, double two, double many, double base)
{
    if (o == 1)
        return q1;
    if (o == 2)
        return two;
    return cv_mul(many, pow(base, (double)(o - 3)));
}

/* weight numerator comb[s] * (1.0 - exp(-lam)), models.py:87 / :221 */
CV_HD double cv_class_mass(double comb_s, double lam)
{
    return cv_mul(comb_s, cv_one_minus_exp_neg(lam));
}

/* Everything of a term that does not depend on the bin.
 *   num / total = b(o) * n_os / sum_s n_os = the weight w of the term (models.py:229, :236)
 *   (lgh, lgl)  = log(lam) as a double-double
 * D(lam) is split as e^lin * Dred with lin exact, so that the scaled value of the term at bin j is
 *   exp(j*log(lam) - lgamma(j+1) + 420 - lin) * f,   f = w / Dred
 * -- one exp per seed and no logarithm per term. */
CV_HD CvTerm cv_term_make(double lam, double num, double total, double lgh, double lgl)
{
    CvTerm t;
    t.lam = 1.0;
    t.lh = t.ll = 0.0;
    t.lin = 0.0;
    if (lam != lam || num != num || total != total) {
        t.f = NAN;
        return t;
    }
    if (!(lam > 0.0) || !(num > 0.0)) { /* zero weight: contributes exactly 0 (a_os * tp = 0) */
        t.f = 0.0;
        return t;
    }
    /* c:25-28: staged reduction by 200 */
    double n = 0.0;
    double r = lam;
    if (lam > 200.0) {
        n = ceil(lam * (1.0 / 200.0)) - 1.0;
        r = cv_fma(-200.0, n, lam); /* exact */
        while (r > 200.0) {
            r = cv_sub(r, 200.0);
            n += 1.0;
        }
        while (r <= 0.0 && n > 0.0) {
            r = cv_add(r, 200.0);
            n -= 1.0;
        }
    }
    double dred;
    if (r > 1e-8 && r < 0x1p-11) {
        /* c:30 `expl(l) - 1` in x87 long double: e^r lies in [1, 2), where the 64-bit format has
         * a spacing of 2^-63, so the difference is e^r - 1 rounded to a multiple of 2^-63 -- a
         * relative perturbation of up to 5e-12 at r = 1e-8 that the high error classes carry
         * straight into p_1.  expm1(r) = r + tail, tail = r^2/2 + ... + r^5/120 (r^6/720 is
         * below 2^-63 * 1e-3 here); both parts are split into integer and fraction of 2^-63. */
        double tail = cv_mul(cv_mul(r, r),
                             cv_fma(r, cv_fma(r, cv_fma(r, 1.0 / 120.0, 1.0 / 24.0), 1.0 / 6.0), 0.5));
        double x1 = cv_mul(r, 0x1p63), x2 = cv_mul(tail, 0x1p63);
        double i1 = floor(x1), i2 = floor(x2);
        double units = cv_add(cv_add(i1, i2), rint(cv_add(cv_sub(x1, i1), cv_sub(x2, i2))));
        t.lin = 200.0 * n;
        dred = cv_mul(units, 0x1p-63);
    } else if (r > 1e-8) { /* c:29-31: D = e^(200 n) (e^r - 1) = e^lam (1 - e^-r) */
        t.lin = lam;
        dred = (r >= 38.0) ? 1.0 : -expm1(-r);
    } else { /* c:19: the ORIGINAL rate is the denominator */
        t.lin = 200.0 * n;
        dred = lam;
    }
    t.lam = lam;
    t.lh = lgh;
    t.ll = lgl;
    t.f = cv_div(num, cv_mul(total, dred));
    return t;
}

/* Scaled value of a term at the head bin j0 of a row:
 *   exp(j0 * log(lam) + [CV_SCALE_LOG - lgamma(j0+1)] - lin) * f
 * (head_h, head_l) is the bracketed row constant as a double-double. */
CV_HD double cv_seed(double j0, double head_h, double head_l, double lh, double ll, double lin,
                     double f)
{
    double p = cv_mul(j0, lh);
    double pe = cv_fma(j0, ll, cv_fma(j0, lh, -p));
    cv_dd s1 = cv_two_sum(p, head_h);
    cv_dd s2 = cv_two_sum(s1.hi, -lin);
    double lo = cv_add(cv_add(pe, head_l), cv_add(s1.lo, s2.lo));
    double eh = cv_add(s2.hi, lo);
    double el = cv_sub(lo, cv_sub(eh, s2.hi));
    double u = exp(eh);
    return cv_mul(cv_fma(u, el, u), f);
}

/* utils.py:32-35 safe_log of a probability that arrives scaled by 2^128, with the classes above */
CV_HD double cv_log_scaled(double ps)
{
    if (ps >= CV_P_BAND)
        return cv_sub(log(ps), CV_PSCALE_LOG);
    if (ps != ps)
        return ps;
    return ps >= CV_P_ZERO ? CV_BAND_LOG : -INFINITY;
}

/* models.py:103-107 for the last step: total mass, tail term */
CV_HD double cv_finish_loglik(double weighted_logs, double mass, double tail)
{
    double tail_term = 0.0;
    if (mass < 1.0)
        tail_term = tail * log(1.0 - mass); /* 1 - mass > 0 here; utils.py:32-35 */
    return weighted_logs + tail_term;
}
