/* Test-only host build of covest_b200/csrc/cvmath.h (g++/gcc, -ffp-contract=off).  It lets the
 * CPU test-suite check the scalar building blocks of the CUDA kernels -- exp() emulation against
 * the container's libm, the double-double log, integer powers -- without a GPU.  Not part of the
 * product: nothing under covest_b200/ loads this. */
#include "../../covest_b200/csrc/cvmath.h"

extern "C" double probe_exp_libm(double x) { return cv_exp_libm(x); }
extern "C" double probe_one_minus_exp_neg(double x) { return cv_one_minus_exp_neg(x); }
extern "C" double probe_pow_uint(double x, int n) { return cv_pow_uint(x, n); }
extern "C" void probe_log_dd(double x, double *hi, double *lo)
{
    cv_dd r = cv_log_dd(x);
    *hi = r.hi;
    *lo = r.lo;
}
/* counts arguments in [lo, hi) (uniform or log-uniform in |x|) whose emulated exp() differs from
 * libm's in any bit */
extern "C" long probe_exp_mismatches(double lo, double hi, long n, int log_uniform, unsigned long long seed)
{
    unsigned long long s = seed ? seed : 88172645463325252ULL;
    long bad = 0;
    for (long i = 0; i < n; i++) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
        double x = log_uniform ? -exp(log(-hi) + u * (log(-lo) - log(-hi))) : lo + u * (hi - lo);
        if (cv_bits(cv_exp_libm(x)) != cv_bits(exp(x)))
            bad++;
    }
    return bad;
}
/* largest |cv_log_tab(x) - logl(x)| / (1 + |logl(x)|) over n log-uniform arguments in [lo, hi) */
extern "C" double probe_log_tab_worst(double lo, double hi, long n, unsigned long long seed)
{
    static double tab[2 * CV_LOG_N];
    cv_log_table(tab);
    unsigned long long s = seed ? seed : 88172645463325252ULL;
    double worst = 0.0;
    for (long i = 0; i < n; i++) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
        double x = exp(log(lo) + u * (log(hi) - log(lo)));
        long double want = logl((long double)x);
        double err = (double)(fabsl((long double)cv_log_tab(x, tab) - want) / (1.0L + fabsl(want)));
        if (err > worst)
            worst = err;
    }
    return worst;
}
extern "C" double probe_log_tab(double x)
{
    static double tab[2 * CV_LOG_N];
    cv_log_table(tab);
    return cv_log_tab(x, tab);
}
