/*
 * faithful.cu -- term-by-term re-evaluation of the points whose result hinges on the reference's
 * roundings to the subnormal grid.
 *
 * The fast kernels (kernels.cu, factored.cu) sum the mixture terms of a bin in scaled arithmetic and
 * round once.  The reference rounds where its Python / C code rounds: every truncated_poisson value
 * is cast to double (c_src/covest_poissonmodule.c:33), every a_os * tp and b(o) * sum product is a
 * double multiplication (covest/models.py:236-239).  Those roundings are invisible (2^-53 relative)
 * unless the probability of a bin with a count is so small that the intermediates are subnormal:
 * then one unit of the subnormal grid is ln 2 or more in log p_j.  The fast kernels mark such
 * points (cvmodel.h: a counted bin with p in the band makes the value < CV_BAND_LL) -- about one
 * point in a hundred of a wide candidate box, the ones whose last counted bins run out of the
 * double range -- and this file evaluates them again the reference's way:
 *
 *   tp(L, j)  e^-lin * prod_{i<=j} L / i (c:22-24) kept as mantissa x 2^e (the reference keeps the
 *             product in an x87 long double, whose range a double mantissa with a separate exponent
 *             covers), times 1 / Dred with the reference's denominator as implemented
 *             (D = e^lin Dred, cv_term_make), cast to double ONCE -- into the subnormal grid when it
 *             is that small (c:33)
 *   a_os * tp, their sum over s, b(o) * sum, the sum over o: IEEE double operations, as Python does
 *             them (sums of subnormal numbers are exact, so their order does not matter)
 *   models.py:100-107 with libm-grade log and a compensated mass
 *
 * One warp per marked point (drawn from a list that a scan of the values fills): the lanes split
 * the bins -- lane l walks the 8 consecutive bins l * 8 + 1 .. of a panel of 256, starting from the
 * product up to its first bin, which a warp-wide multiplicative scan of the lanes' own partial
 * products supplies -- and every lane runs over all (copy number, error class) terms, so the sums
 * over s and o are plain sequential sums in registers.  Only bins up to the last one with a count
 * are needed when the histogram has no tail (models.py:103-104: the mass then does not enter).
 * A few microseconds per point with two copy numbers; points with many copy numbers (their cost grows
 * with copies x bins) are worked off by a whole CTA each, the copy numbers dealt to its warps.  The
 * kernels are correct for any point (path mode 5 runs every point through them,
 * tests/test_gpu_big_golden.py).
 *
 * Reference lines are relative to /root/reference.
 */
#include "faithful.h"

#include "kdevice.h"

#define CVX_THREADS 256
#define CVX_KB 8 /* bins per lane and panel */
#define CVX_NP 2 /* passes of 32 error classes (CV_MAX_ERR = 64) */

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

/* mantissa into [1, 2), the exponent it had into e (m positive and normal; 0, inf, NaN pass) */
__device__ __forceinline__ void cvx_norm(double &m, int &e)
{
    const int hi = __double2hiint(m);
    const int ex = ((hi >> 20) & 0x7ff);
    if (ex == 0 || ex == 0x7ff)
        return;
    m = __hiloint2double(hi - ((ex - 1023) << 20), __double2loint(m));
    e += ex - 1023;
}

/* x * 2^e rounded to double ONCE (round to nearest even), also into the subnormal grid; x >= 0 */
__device__ __forceinline__ double cvx_scale_once(double x, int e)
{
    const int hi = __double2hiint(x);
    const int ex = (hi >> 20) & 0x7ff;
    if (ex == 0 || ex == 0x7ff) /* 0 (x is never subnormal here), inf, NaN */
        return x;
    const int E = ex - 1023 + e; /* exponent of the result */
    if (E > 1023)
        return INFINITY;
    if (E >= -1022)
        return __hiloint2double(hi + (e << 20), __double2loint(x));
    if (E < -1076)
        return 0.0;
    /* units of 2^-1074: t = x * 2^(e + 1074) has exponent E + 1074 in [-2, 51] */
    const double t = __hiloint2double(hi + ((e + 1074) << 20), __double2loint(x));
    return __longlong_as_double((long long)rint(t)); /* the bits of a subnormal are its units */
}

struct CvxPoint {
    double c, e, q1, two, many, base;
    int n_copies;
};

/* The bin probabilities of one point, summed over the copy numbers o_first, o_first + o_step, ...:
 * p_j at acc[j - 1] (ft.acc_doubles doubles of scratch; every lane touches its own bins only). */
__device__ void cvx_accumulate(int lane, const CvModelDesc &m, const CvxPoint &P, const CvFaithTables &ft,
                               double *acc, int o_first, int o_step)
{
    const int S = m.n_err;
    const int J = m.tail != 0.0 ? ft.j_all : ft.j_counted; /* bins that enter the result */
    const int npass = (S + 31) >> 5;
    const double ck = cv_kmer_coverage(P.c, m.k, m.r);
    for (int j = lane; j < J; j += 32)
        acc[j] = 0.0;
    __syncwarp();
    for (int o = o_first; o <= P.n_copies; o += o_step) {
        const double b_o = m.model_kind ? cv_copy_weight(o, P.q1, P.two, P.many, P.base) : 1.0;
        /* models.py:221-232: n_os = comb[s] * (1.0 - exp(o * -l_s)), a_os = n_os / (sum_s n_os or 1),
         * the sum left to right as Python's.  Lane l holds the classes l and l + 32. */
        double my_lam[CVX_NP], my_nos[CVX_NP], my_a[CVX_NP], my_f[CVX_NP], my_cm[CVX_NP];
        int my_ce[CVX_NP];
        double tot = 0.0;
#pragma unroll
        for (int u = 0; u < CVX_NP; u++) {
            const int s = 32 * u + lane;
            my_lam[u] = my_nos[u] = 0.0;
            if (s < S) {
                my_lam[u] = cv_mul((double)o, cv_error_class_rate(ck, m.pow3[s], P.e, m.k, s));
                my_nos[u] = cv_class_mass(m.comb[s], my_lam[u]);
            }
            if (u < npass)
                for (int i = 0; i < min(32, S - 32 * u); i++) {
                    const double v = __shfl_sync(CV_FULL_MASK, my_nos[u], i);
                    tot = (u == 0 && i == 0) ? cv_add(0.0, v) : cv_add(tot, v);
                }
        }
        if (tot == 0.0)
            tot = 1.0; /* utils.py:25-29 fix_zero */
#pragma unroll
        for (int u = 0; u < CVX_NP; u++) {
            my_a[u] = cv_div(my_nos[u], tot);
            my_f[u] = 0.0;
            my_cm[u] = 1.0;
            my_ce[u] = 0;
            if (32 * u + lane < S) {
                const CvTerm t = cv_term_make(my_lam[u], 1.0, 1.0, 0.0, 0.0); /* f = 1 / Dred, D = e^lin Dred */
                my_f[u] = t.f;
                if (t.f != 0.0 && t.f == t.f) /* the carry: the product up to the panel's first bin */
                    cvx_exp_neg(t.lin, my_cm[u], my_ce[u]);
            }
        }
        for (int j0 = 0; j0 < J; j0 += 32 * CVX_KB) { /* panels of 256 bins */
            const int jl = j0 + lane * CVX_KB; /* the lane walks bins jl + 1 .. jl + CVX_KB */
            double inner[CVX_KB], rc[CVX_KB];
#pragma unroll
            for (int k = 0; k < CVX_KB; k++) {
                inner[k] = 0.0;
                rc[k] = jl + k < J ? __ldg(ft.rcp + jl + k) : 1.0; /* 1 / j of the lane's bins */
            }
#pragma unroll
            for (int u = 0; u < CVX_NP; u++) {
                if (u >= npass)
                    break;
                const int sn = min(32, S - 32 * u);
                for (int s = 0; s < sn; s++) {
                    const double lam = __shfl_sync(CV_FULL_MASK, my_lam[u], s);
                    const double a = __shfl_sync(CV_FULL_MASK, my_a[u], s);
                    const double f = __shfl_sync(CV_FULL_MASK, my_f[u], s);
                    double cm = __shfl_sync(CV_FULL_MASK, my_cm[u], s);
                    int ce = __shfl_sync(CV_FULL_MASK, my_ce[u], s);
                    if (f == 0.0 || a == 0.0) /* a_os * tp = 0 exactly (zero rate or zero weight; tp is finite) */
                        continue;
                    /* the lane's own partial product, then the products of the lanes before it */
                    double pm = 1.0;
                    int pe = 0;
#pragma unroll
                    for (int k = 0; k < CVX_KB; k++) {
                        pm = cv_mul(pm, cv_mul(lam, rc[k])); /* c:22-24 */
                        if (k == 3 || k == CVX_KB - 1)
                            cvx_norm(pm, pe);
                    }
                    double sm = pm; /* inclusive scan over the lanes */
                    int se = pe;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double om = __shfl_up_sync(CV_FULL_MASK, sm, d);
                        const int oe = __shfl_up_sync(CV_FULL_MASK, se, d);
                        if (lane >= d) {
                            sm = cv_mul(sm, om);
                            se += oe;
                            cvx_norm(sm, se);
                        }
                    }
                    double xm = __shfl_up_sync(CV_FULL_MASK, sm, 1); /* exclusive */
                    int xe = __shfl_up_sync(CV_FULL_MASK, se, 1);
                    if (lane == 0) {
                        xm = 1.0;
                        xe = 0;
                    }
                    double vm = cv_mul(cm, xm);
                    int ve = ce + xe;
                    cvx_norm(vm, ve);
#pragma unroll
                    for (int k = 0; k < CVX_KB; k++) {
                        vm = cv_mul(vm, cv_mul(lam, rc[k]));
                        if (k == 3)
                            cvx_norm(vm, ve);
                        /* c:33: the quotient, cast to double once; models.py:236-238: times a_os, summed */
                        const double tp = cvx_scale_once(cv_mul(vm, f), ve);
                        inner[k] = cv_add(inner[k], cv_mul(a, tp));
                    }
                    /* the carry moves to the end of the panel */
                    const double tm = __shfl_sync(CV_FULL_MASK, sm, 31);
                    const int te = __shfl_sync(CV_FULL_MASK, se, 31);
                    cm = cv_mul(cm, tm);
                    ce += te;
                    cvx_norm(cm, ce);
                    if (lane == s) {
                        my_cm[u] = cm;
                        my_ce[u] = ce;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < CVX_KB; k++)
                if (jl + k < J) /* models.py:236: b_o(o) * sum(...), summed over o */
                    acc[jl + k] = cv_add(acc[jl + k], cv_mul(b_o, inner[k]));
        }
        __syncwarp();
    }
}

/* models.py:100-107 from the bin probabilities, which are the sums of n_acc arrays `stride` doubles
 * apart (the shares of the warps that split the copy numbers), added in a fixed order */
__device__ double cvx_finish(int lane, const CvModelDesc &m, const CvFaithTables &ft, const double *acc, int n_acc,
                             long long stride)
{
    const int J = m.tail != 0.0 ? ft.j_all : ft.j_counted;
    CvPartial part;
    part.sum = 0.0;
    part.mass_h = part.mass_l = 0.0;
    for (int j = lane; j < J; j += 32) {
        const double h = __ldg(ft.cnt_of_j + j); /* < 0: j + 1 is not a key of hist */
        if (h < 0.0)
            continue;
        double p = acc[j];
        for (int a = 1; a < n_acc; a++)
            p = cv_add(p, acc[a * stride + j]);
        cv_partial_add_mass(part, p);
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

__device__ __forceinline__ CvxPoint cvx_load_point(const CvModelDesc &m, const CvLattice &lat,
                                                   const double *__restrict__ params, long long pi, int clip)
{
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
    return P;
}

/* the marked points of a batch: their indices, in any order */
__global__ void __launch_bounds__(256)
cv_marked_list_kernel(const double *__restrict__ out_ll, long long n, unsigned int *__restrict__ list,
                      unsigned long long *__restrict__ count)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double v = i < n ? out_ll[i] : 0.0;
    const bool marked = i < n && v < CV_BAND_LL && v > -INFINITY;
    const unsigned int ballot = __ballot_sync(CV_FULL_MASK, marked);
    if (ballot == 0)
        return;
    const int lane = threadIdx.x & 31;
    unsigned long long at = 0;
    if (lane == 0)
        at = atomicAdd(count, (unsigned long long)__popc(ballot));
    at = __shfl_sync(CV_FULL_MASK, at, 0);
    if (marked)
        list[at + __popc(ballot & ((1u << lane) - 1))] = (unsigned int)i;
}

/* One warp per marked point.  Points with more than CVX_HEAVY copy numbers would keep a single warp
 * busy for milliseconds (cfg4: 50 copies x 5000 bins): they go to a second list, which
 * cv_faithful_heavy_kernel works off with a whole CTA per point. */
#define CVX_HEAVY 8
#define CVX_HEAVY_THREADS 512
__global__ void __launch_bounds__(CVX_THREADS)
cv_faithful_kernel(const __grid_constant__ CvModelDesc m, const __grid_constant__ CvLattice lat,
                   const double *__restrict__ params, int clip, double *__restrict__ out_ll, CvFaithTables ft,
                   const unsigned int *__restrict__ list, unsigned long long *__restrict__ counters,
                   unsigned int *__restrict__ heavy_list)
{
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double *acc = ft.scratch + gw * ft.acc_doubles;
    const unsigned long long total = counters[0];
    for (;;) {
        unsigned long long it = 0;
        if (lane == 0)
            it = atomicAdd(counters + 1, 1ULL);
        it = __shfl_sync(CV_FULL_MASK, it, 0);
        if (it >= total)
            break;
        const long long pi = list[it];
        const CvxPoint P = cvx_load_point(m, lat, params, pi, clip);
        if (P.n_copies > CVX_HEAVY) {
            if (lane == 0)
                heavy_list[atomicAdd(counters + 2, 1ULL)] = (unsigned int)pi;
            continue;
        }
        cvx_accumulate(lane, m, P, ft, acc, 1, 1);
        const double r = cvx_finish(lane, m, ft, acc, 1, 0);
        if (lane == 0)
            out_ll[pi] = r;
        __syncwarp();
    }
}

/* One CTA per point with many copy numbers: warp w takes the copies w + 1, w + 1 + W, ..., the shares
 * meet in cvx_finish in a fixed order (the value does not depend on which CTA got the point). */
__global__ void __launch_bounds__(CVX_HEAVY_THREADS, 1)
cv_faithful_heavy_kernel(const __grid_constant__ CvModelDesc m, const __grid_constant__ CvLattice lat,
                         const double *__restrict__ params, int clip, double *__restrict__ out_ll, CvFaithTables ft,
                         const unsigned int *__restrict__ heavy_list, unsigned long long *__restrict__ counters)
{
    __shared__ unsigned long long s_it;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = CVX_HEAVY_THREADS / 32;
    double *acc0 = ft.scratch + (long long)blockIdx.x * W * ft.acc_doubles;
    const unsigned long long total = counters[2];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0)
            s_it = atomicAdd(counters + 3, 1ULL);
        __syncthreads();
        const unsigned long long it = s_it;
        if (it >= total)
            break;
        const long long pi = heavy_list[it];
        const CvxPoint P = cvx_load_point(m, lat, params, pi, clip);
        cvx_accumulate(lane, m, P, ft, acc0 + warp * ft.acc_doubles, warp + 1, W);
        __threadfence_block();
        __syncthreads();
        if (warp == 0) {
            const double r = cvx_finish(lane, m, ft, acc0, P.n_copies < W ? P.n_copies : W, ft.acc_doubles);
            if (lane == 0)
                out_ll[pi] = r;
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
                               int clip, double *out_ll, const CvFaithTables &ft, int n_sm, unsigned int *list,
                               unsigned long long *counters, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    /* counters: marked points, cursor, points with many copies, cursor */
    cudaError_t e = cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess)
        return e;
    cv_marked_list_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, stream>>>(out_ll, n, list, counters);
    if ((e = cudaGetLastError()) != cudaSuccess)
        return e;
    long long ctas = (n + (CVX_THREADS / 32) - 1) / (CVX_THREADS / 32); /* never more warps than points */
    if (ctas > 2 * n_sm)
        ctas = 2 * n_sm;
    cv_faithful_kernel<<<(unsigned int)ctas, CVX_THREADS, 0, stream>>>(m, lat, params, clip, out_ll, ft, list, counters,
                                                                   list + n);
    if ((e = cudaGetLastError()) != cudaSuccess)
        return e;
    /* the same scratch: n_sm CTAs x 16 warps = the 2 n_sm x 8 warps of the kernel before */
    cv_faithful_heavy_kernel<<<(unsigned int)(n < n_sm ? n : n_sm), CVX_HEAVY_THREADS, 0, stream>>>(
        m, lat, params, clip, out_ll, ft, list + n, counters);
    return cudaGetLastError();
}
