/*
 * faithful.cu -- term-by-term re-evaluation of the few points whose result hinges on the
 * reference's roundings to the subnormal grid.
 *
 * The fast kernels (kernels.cu, factored.cu) sum the mixture terms of a bin in scaled arithmetic and
 * round once.  The reference rounds where its Python / C code rounds: every truncated_poisson value
 * is cast to double (c_src/covest_poissonmodule.c:33), every a_os * tp and b(o) * sum product is a
 * double multiplication (covest/models.py:236-239).  Those roundings are invisible (2^-53 relative)
 * unless the probability of a bin with a count is so small that the intermediates are subnormal:
 * then one unit of the subnormal grid is ln 2 or more in log p_j.  The fast kernels mark such
 * points (cvmodel.h, CV_BAND_LOG: a counted bin with 2^-1088 <= p < 2^-1000 makes the value
 * < CV_BAND_LL), and this kernel evaluates them again the reference's way:
 *
 *   tp(L, j)  the running product prod_{i<=j} L / i of c:22-24 with the quotient a double division
 *             as in the reference, the product kept as mantissa x 2^e (the reference keeps it in an
 *             x87 long double, whose range a double mantissa with a separate exponent covers),
 *             times 1 / D(L) with the reference's denominator as implemented (cv_term_make), cast
 *             to double ONCE -- into the subnormal grid when it is that small (c:33)
 *   a_os * tp, their sum over s, b(o) * sum, the sum over o: IEEE double operations, as Python does
 *             them (sums of subnormal numbers are exact, so their order does not matter)
 *   models.py:100-107 with libm-grade log and a compensated mass
 *
 * One warp per marked point: its lanes are the (copy number, error class) terms of a pass of
 * copies, all walking the bins together.  Marked points have few copies (every term must be tiny at
 * a bin that has a count), so this costs microseconds per batch; it is correct for any point.
 *
 * Reference lines are relative to /root/reference.
 */
#include "faithful.h"

#include "kdevice.h"

#define CVX_THREADS 128

/* exp(-lin), lin >= 0, as mant * 2^e2 with mant in [1/2, 2) */
__device__ __forceinline__ void cvx_exp_neg(double lin, double &mant, int &e2)
{
    const double kf = rint(lin * -0x1.71547652b82fep+0); /* -lin / ln 2 */
    /* -lin - kf ln2 with ln2 = hi + lo, kf * hi exact (hi has 33 bits, |kf| < 2^20) */
    double r = cv_fma(-kf, 0x1.62e42fee00000p-1, -lin);
    r = cv_fma(-kf, 0x1.a39ef35793c76p-33, r);
    mant = exp(r);
    e2 = (int)kf;
}

struct CvxPoint {
    double c, e, q1, two, many, base;
    int n_copies;
};

/* The log-likelihood of one point, evaluated by the whole warp.  pacc: n_bins doubles of scratch. */
__device__ double cvx_point(int lane, const CvModelDesc &m, const CvxPoint &P, const CvFaithTables &ft,
                            double *pacc)
{
    const int S = m.n_err;
    /* lanes of a copy: the error classes, rounded up to a power of two; classes beyond 32 share lanes */
    int sp = 1;
    while (sp < S && sp < 32)
        sp <<= 1;
    const int ns = (S + sp - 1) / sp; /* classes per lane: 1, or 2 for 33..64 classes */
    const int cpt = 32 / sp;          /* copies per pass */
    const int ls = lane & (sp - 1), lc = lane / sp;
    const double ck = cv_kmer_coverage(P.c, m.k, m.r);
    double l_s[2] = {0.0, 0.0};
    double comb[2] = {0.0, 0.0};
    for (int u = 0; u < ns; u++) {
        const int s = ls + sp * u;
        if (s < S) {
            l_s[u] = cv_error_class_rate(ck, m.pow3[s], P.e, m.k, s);
            comb[u] = m.comb[s];
        }
    }
    for (int b = lane; b < ft.n; b += 32)
        pacc[b] = 0.0;
    __syncwarp();
    for (int o0 = 1; o0 <= P.n_copies; o0 += cpt) {
        const int o = o0 + lc;
        const bool live = o <= P.n_copies;
        /* models.py:221-232: n_os = comb[s] * (1.0 - exp(o * -l_s)), a_os = n_os / (sum_s n_os or 1) */
        double lam[2], nos[2];
        for (int u = 0; u < 2; u++) {
            const bool on = live && u < ns && ls + sp * u < S;
            lam[u] = on ? cv_mul((double)o, l_s[u]) : 0.0;
            nos[u] = on ? cv_class_mass(comb[u], lam[u]) : 0.0;
        }
        double tot = 0.0; /* left to right over s, as Python's sum */
        for (int s = 0; s < S; s++) {
            const double v = __shfl_sync(CV_FULL_MASK, s < sp ? nos[0] : nos[1], (lane & ~(sp - 1)) + (s & (sp - 1)));
            tot = s == 0 ? cv_add(0.0, v) : cv_add(tot, v);
        }
        if (tot == 0.0)
            tot = 1.0; /* utils.py:25-29 fix_zero */
        const double b_o = !live ? 0.0 : m.model_kind ? cv_copy_weight(o, P.q1, P.two, P.many, P.base) : 1.0;
        double a[2], f[2], mant[2];
        int e2[2];
        for (int u = 0; u < 2; u++) {
            a[u] = cv_div(nos[u], tot);
            const CvTerm t = cv_term_make(lam[u], 1.0, 1.0, 0.0, 0.0); /* f = 1 / Dred, D = e^lin Dred */
            f[u] = t.f;
            mant[u] = 1.0;
            e2[u] = 0;
            if (t.f != 0.0 && t.f == t.f)
                cvx_exp_neg(t.lin, mant[u], e2[u]);
        }
        int jprev = 0;
        for (int b = 0; b < ft.n; b++) {
            const int j = ft.key[b];
            double inner = 0.0;
            for (int u = 0; u < 2; u++) {
                if (u >= ns)
                    break;
                for (int i = jprev + 1; i <= j; i++) { /* c:22-24 */
                    mant[u] = cv_mul(mant[u], cv_div(lam[u], (double)i));
                    if (mant[u] < 0x1p-400) {
                        mant[u] = cv_mul(mant[u], 0x1p400);
                        e2[u] -= 400;
                    } else if (mant[u] > 0x1p400) {
                        mant[u] = cv_mul(mant[u], 0x1p-400);
                        e2[u] += 400;
                    }
                }
                /* c:33: the quotient, cast to double once (ldexp rounds once, also into the
                 * subnormal grid) */
                double tp = 0.0;
                if (f[u] != 0.0) {
                    int ex; /* mantissa in [1/2, 1): the product with 1 / Dred stays a normal number */
                    const double v = cv_mul(frexp(mant[u], &ex), f[u]);
                    const int ee = e2[u] + ex;
                    tp = ldexp(v, ee < -4000 ? -4000 : ee > 4000 ? 4000 : ee);
                }
                inner = cv_add(inner, cv_mul(a[u], tp)); /* models.py:236-238 */
            }
            jprev = j;
            for (int d = 1; d < sp; d <<= 1)
                inner = cv_add(inner, __shfl_xor_sync(CV_FULL_MASK, inner, d));
            double outer = cv_mul(b_o, inner); /* models.py:236 b_o(o) * sum(...) */
            if (!live)
                outer = 0.0;
            for (int d = sp; d < 32; d <<= 1)
                outer = cv_add(outer, __shfl_xor_sync(CV_FULL_MASK, outer, d));
            if (lane == 0)
                pacc[b] = cv_add(pacc[b], outer);
        }
        __syncwarp();
    }
    /* models.py:100-107 */
    CvPartial part;
    part.sum = 0.0;
    part.mass_h = part.mass_l = 0.0;
    for (int b = lane; b < ft.n; b += 32) {
        const double p = pacc[b];
        cv_partial_add_mass(part, p);
        const double h = ft.cnt[b];
        if (h != 0.0)
            part.sum = cv_add(part.sum, cv_mul(h, p <= 0.0 ? -INFINITY : log(p))); /* utils.py:32-35 */
    }
    for (int d = 16; d >= 1; d >>= 1) {
        CvPartial q;
        q.sum = __shfl_xor_sync(CV_FULL_MASK, part.sum, d);
        q.mass_h = __shfl_xor_sync(CV_FULL_MASK, part.mass_h, d);
        q.mass_l = __shfl_xor_sync(CV_FULL_MASK, part.mass_l, d);
        cv_partial_merge(part, q);
    }
    __syncwarp();
    double mass = cv_add(part.mass_h, part.mass_l);
    if (!(mass < 1.0))
        mass = 1.0;
    return cv_finish_loglik(part.sum, mass, m.tail);
}

__global__ void __launch_bounds__(CVX_THREADS)
cv_faithful_kernel(const __grid_constant__ CvModelDesc m, const __grid_constant__ CvLattice lat,
                   const double *__restrict__ params, long long n, int clip, double *__restrict__ out_ll,
                   CvFaithTables ft, unsigned long long *__restrict__ n_fixed)
{
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    double *pacc = ft.scratch + gw * ft.n;
    for (long long base = gw * 32; base < n; base += nw * 32) {
        const long long i = base + lane;
        const double v = i < n ? out_ll[i] : 0.0;
        unsigned int marked = __ballot_sync(CV_FULL_MASK, i < n && v < CV_BAND_LL && v > -INFINITY);
        while (marked) {
            const int src = __ffs(marked) - 1;
            marked &= marked - 1;
            const long long pi = base + src;
            double row[CV_MAX_PARAMS];
            cvf_raw_row(m, lat, params, pi, row);
            CvxPoint P;
            P.c = cvf_clipped(m, row, clip, 0);
            P.e = cvf_clipped(m, row, clip, 1);
            P.q1 = P.two = P.many = P.base = 0.0;
            P.n_copies = 1; /* basic model: the single copy o = 1 with weight 1 */
            if (m.model_kind) {
                P.q1 = cvf_clipped(m, row, clip, 2);
                const double q2 = cvf_clipped(m, row, clip, 3), q = cvf_clipped(m, row, clip, 4);
                P.two = cv_mul(cv_sub(1.0, P.q1), q2);
                P.many = cv_mul(cv_mul(cv_sub(1.0, P.q1), cv_sub(1.0, q2)), q);
                P.base = cv_sub(1.0, q);
                P.n_copies = cvf_cutoff(m, P.q1, P.two, P.many, P.base) - 1; /* models.py:235 */
            }
            const double r = cvx_point(lane, m, P, ft, pacc);
            if (lane == 0) {
                out_ll[pi] = r;
                if (n_fixed)
                    atomicAdd(n_fixed, 1ULL);
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(256) cv_mark_all_kernel(double *__restrict__ out_ll, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out_ll[i] = 2.0 * CV_BAND_LL;
}

cudaError_t cv_launch_mark_all(double *out_ll, long long n, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    cv_mark_all_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, stream>>>(out_ll, n);
    return cudaGetLastError();
}

int cv_faithful_warps(int n_sm) { return n_sm * 2 * (CVX_THREADS / 32); }

cudaError_t cv_launch_faithful(const CvModelDesc &m, const CvLattice &lat, const double *params, long long n,
                               int clip, double *out_ll, const CvFaithTables &ft, int n_sm,
                               unsigned long long *n_fixed, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    long long ctas = (n + CVX_THREADS - 1) / CVX_THREADS;
    if (ctas > 2 * n_sm)
        ctas = 2 * n_sm;
    cv_faithful_kernel<<<(unsigned int)ctas, CVX_THREADS, 0, stream>>>(m, lat, params, n, clip, out_ll, ft, n_fixed);
    return cudaGetLastError();
}
