/*
 * topk.cu -- K3 for large batches: the K best log-likelihoods by one stable device radix sort.
 *
 * Total order (the same as cv_topk_select, kernels.cu): larger value first, ties to the lower
 * index, NaN counts as -inf.  The values are mapped to order-preserving 64-bit keys and sorted in
 * descending order together with their indices (cub, stable: equal keys keep ascending index);
 * the first K pairs are the selection.  A full sort of 10^6 pairs costs ~0.2 ms on a B200, a
 * seventh of the K-pass selection kernel it replaces at that size.
 */
#include <cub/device/device_radix_sort.cuh>

#include "kernels.h"

__global__ void __launch_bounds__(256)
cv_topk_keys(const double *__restrict__ vals, long long n, unsigned long long *__restrict__ keys,
             unsigned int *__restrict__ idx)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    double v = vals[i];
    if (v != v)
        v = -INFINITY;
    if (v == 0.0)
        v = 0.0; /* -0.0 and +0.0 tie */
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    keys[i] = (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
    idx[i] = (unsigned int)i;
}

__global__ void cv_topk_take(const unsigned long long *__restrict__ keys, const unsigned int *__restrict__ idx,
                             long long n, int K, double *__restrict__ out_v, long long *__restrict__ out_i)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K)
        return;
    if (k >= n) {
        out_v[k] = -INFINITY;
        out_i[k] = -1;
        return;
    }
    unsigned long long key = keys[k];
    unsigned long long b = (key >> 63) ? (key & 0x7fffffffffffffffULL) : ~key;
    out_v[k] = __longlong_as_double((long long)b);
    out_i[k] = (long long)idx[k];
}

size_t cv_topk_sort_bytes(long long n)
{
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, (unsigned long long *)nullptr,
                                              (unsigned long long *)nullptr, (unsigned int *)nullptr,
                                              (unsigned int *)nullptr, (int)n, 0, 64, (cudaStream_t)0);
    const size_t a = ((size_t)n * 8 + 255) & ~(size_t)255, b = ((size_t)n * 4 + 255) & ~(size_t)255;
    return 2 * a + 2 * b + tmp + 256;
}

cudaError_t cv_launch_topk_sort(const double *ll, long long n, int K, void *scratch, size_t scratch_bytes,
                                double *out_ll, long long *out_idx, cudaStream_t stream)
{
    if (n > 0x7fffffffLL)
        return cudaErrorInvalidValue;
    const size_t a = ((size_t)n * 8 + 255) & ~(size_t)255, b = ((size_t)n * 4 + 255) & ~(size_t)255;
    unsigned char *base = (unsigned char *)scratch;
    unsigned long long *k0 = (unsigned long long *)base, *k1 = (unsigned long long *)(base + a);
    unsigned int *i0 = (unsigned int *)(base + 2 * a), *i1 = (unsigned int *)(base + 2 * a + b);
    void *tmp = base + 2 * a + 2 * b;
    size_t tmp_bytes = scratch_bytes - (2 * a + 2 * b);
    if (n > 0) {
        cv_topk_keys<<<(unsigned int)((n + 255) / 256), 256, 0, stream>>>(ll, n, k0, i0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    cub::DoubleBuffer<unsigned long long> dk(k0, k1);
    cub::DoubleBuffer<unsigned int> dv(i0, i1);
    if (n > 0) {
        cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, dk, dv, (int)n, 0, 64, stream);
        if (e != cudaSuccess)
            return e;
    }
    cv_topk_take<<<(K + 127) / 128, 128, 0, stream>>>(dk.Current(), dv.Current(), n, K, out_ll, out_idx);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------- */
/* Merge of the best rows of several ranks (multi-GPU rounds): n rows of c doubles, column 0 the   */
/* log-likelihood.  Total order: larger log-likelihood first (NaN as -inf), ties by the parameter */
/* columns in ascending lexicographic order (NaN last) -- for a lattice with ascending axes the   */
/* lattice index, so the result does not depend on how many ranks produced the rows.  One CTA,    */
/* bitonic sort of row indices in shared memory; n <= CV_MERGE_MAX.                                */
/* ------------------------------------------------------------------------------------------- */
#define CV_MERGE_MAX 2048

__device__ __forceinline__ bool cv_row_before(const double *__restrict__ rows, int c, int a, int b)
{
    if (a < 0 || b < 0)
        return b < 0 && a >= 0; /* padding sorts last */
    double x = rows[(size_t)a * c], y = rows[(size_t)b * c];
    if (x != x)
        x = -INFINITY;
    if (y != y)
        y = -INFINITY;
    if (x != y)
        return x > y;
    for (int j = 1; j < c; j++) {
        double u = rows[(size_t)a * c + j], v = rows[(size_t)b * c + j];
        if (u != u)
            u = INFINITY;
        if (v != v)
            v = INFINITY;
        if (u != v)
            return u < v;
    }
    return a < b; /* identical rows: stable */
}

__global__ void __launch_bounds__(1024)
cv_merge_rows_kernel(const double *__restrict__ rows, int n, int c, int k, double *__restrict__ out)
{
    __shared__ int idx[CV_MERGE_MAX];
    int m = 1;
    while (m < n)
        m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        idx[i] = i < n ? i : -1;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const int a = idx[i], b = idx[j];
                    if (cv_row_before(rows, c, b, a) == up) {
                        idx[i] = b;
                        idx[j] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (int t = threadIdx.x; t < k * c; t += blockDim.x) {
        const int r = t / c, j = t - r * c;
        out[t] = (r < n) ? rows[(size_t)idx[r] * c + j] : (j == 0 ? -INFINITY : NAN);
    }
}

cudaError_t cv_launch_merge_rows(const double *rows, int n, int c, int k, double *out, cudaStream_t stream)
{
    if (n < 0 || n > CV_MERGE_MAX || c < 1 || k < 1)
        return cudaErrorInvalidValue;
    cv_merge_rows_kernel<<<1, 1024, 0, stream>>>(rows, n, c, k, out);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------- */
/* K3 for large batches and small K: radix *selection* instead of a full sort.                   */
/*                                                                                               */
/* Every value gets the 96-bit key (order-preserving value bits, ~index): all keys differ, larger */
/* key = earlier in the order of cv_topk_select (larger value first, ties to the lower index, NaN */
/* as -inf).  Rounds of 11 bits from the top: a histogram of the next digit over the keys that     */
/* still match the prefix, then CTA 0 picks the digit that holds the K-th key and extends the      */
/* prefix.  The rounds stop once at most CV_SEL_CAP keys match the prefix; those and everything    */
/* above go to a candidate list that CTA 0 sorts.  Typically three or four passes over the values  */
/* (8 bytes each) instead of the eight passes over 12-byte pairs of the sort, in one launch:       */
/* 0.09 ms instead of 0.22 ms for 10^6 values on a B200.                                           */
/* ------------------------------------------------------------------------------------------- */
#define CV_SEL_BITS 11
#define CV_SEL_BINS (1 << CV_SEL_BITS)
#define CV_SEL_CAP 2048
#define CV_SEL_KMAX 1024
#define CV_SEL_ROUNDS 9 /* 96 bits */

typedef unsigned __int128 cv_u128;

struct CvSelState {
    unsigned long long prefix_hi; /* the key bits fixed so far, right-aligned in (hi:lo) */
    unsigned long long prefix_lo;
    int bits_done;
    int done;
    long long k_rem;              /* keys still to take among those that match the prefix */
    unsigned int n_cand;
    unsigned int pad;
};

__device__ __forceinline__ cv_u128 cv_sel_key(double v, unsigned int i)
{
    if (v != v)
        v = -INFINITY;
    if (v == 0.0)
        v = 0.0;
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    b = (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
    return ((cv_u128)b << 32) | (cv_u128)(unsigned int)(~i);
}

/* The rounds, the collection and the final sort in ONE cooperative launch: the grid (all CTAs
 * resident) meets at grid-wide barriers between the steps; CTA 0 does the picks and the finish. */
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512)
cv_sel_fused(const double *__restrict__ vals, long long n, int K, CvSelState *st, unsigned int *hist,
             unsigned long long *__restrict__ cand_key, unsigned int *__restrict__ cand_idx,
             double *__restrict__ out_v, long long *__restrict__ out_i)
{
    cg::grid_group grid = cg::this_grid();
    volatile CvSelState *vs = st; /* the state changes between barriers: never from a stale L1 line */
    extern __shared__ unsigned char cv_sel_raw[];
    unsigned int *h = reinterpret_cast<unsigned int *>(cv_sel_raw); /* CV_SEL_BINS counters */
    const long long gstride = (long long)gridDim.x * blockDim.x;
    const long long g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < CV_SEL_BINS; i += blockDim.x)
            hist[i] = 0;
        if (threadIdx.x == 0) {
            st->prefix_hi = st->prefix_lo = 0;
            st->bits_done = 0;
            st->done = 0;
            st->k_rem = K;
            st->n_cand = 0;
        }
    }
    grid.sync();
    for (int r = 0; r < CV_SEL_ROUNDS; r++) {
        if (vs->done)
            break; /* the same for every thread of the grid: set before the last barrier */
        /* histogram of the next digit over the keys that match the prefix */
        for (int i = threadIdx.x; i < CV_SEL_BINS; i += blockDim.x)
            h[i] = 0;
        __syncthreads();
        const int bits_done = vs->bits_done;
        const int width = min(CV_SEL_BITS, 96 - bits_done);
        const int shift = 96 - bits_done - width;
        const cv_u128 prefix = ((cv_u128)vs->prefix_hi << 64) | vs->prefix_lo;
        for (long long i = g0; i < n; i += gstride) {
            const cv_u128 key = cv_sel_key(vals[i], (unsigned int)i);
            if (bits_done == 0 || (key >> (96 - bits_done)) == prefix)
                atomicAdd(&h[(unsigned int)(key >> shift) & ((1u << width) - 1u)], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < CV_SEL_BINS; i += blockDim.x)
            if (h[i])
                atomicAdd(&hist[i], h[i]);
        grid.sync();
        if (blockIdx.x == 0) { /* pick the digit that holds the k_rem-th largest matching key */
            unsigned long long *above = reinterpret_cast<unsigned long long *>(h + CV_SEL_BINS);
            for (int i = threadIdx.x; i < CV_SEL_BINS; i += blockDim.x) {
                h[i] = ((volatile unsigned int *)hist)[i];
                hist[i] = 0;
            }
            __syncthreads();
            /* suffix sums: each thread its run of 4 digits, then the runs by one warp */
            unsigned int *runsum = reinterpret_cast<unsigned int *>(above + CV_SEL_BINS);
            {
                const int t = threadIdx.x; /* 512 threads x 4 digits */
                runsum[t] = h[4 * t] + h[4 * t + 1] + h[4 * t + 2] + h[4 * t + 3];
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long run = 0;
                for (int t = 511; t >= 0; t--) {
                    const unsigned int c = runsum[t];
                    runsum[t] = 0;
                    above[4 * t + 3] = run; /* keys above the run's top digit */
                    run += c;
                }
            }
            __syncthreads();
            {
                const int t = threadIdx.x;
                unsigned long long a = above[4 * t + 3];
                for (int d = 4 * t + 3; d >= 4 * t; d--) {
                    above[d] = a;
                    a += h[d];
                }
            }
            __syncthreads();
            const long long k_rem = vs->k_rem;
            for (int d = threadIdx.x; d < CV_SEL_BINS; d += blockDim.x)
                if ((long long)above[d] < k_rem && (long long)(above[d] + h[d]) >= k_rem) {
                    cv_u128 p2 = (prefix << width) | (cv_u128)(unsigned int)d;
                    st->prefix_hi = (unsigned long long)(p2 >> 64);
                    st->prefix_lo = (unsigned long long)p2;
                    st->bits_done = bits_done + width;
                    st->k_rem = k_rem - (long long)above[d];
                    if (h[d] <= CV_SEL_CAP || bits_done + width >= 96)
                        st->done = 1;
                }
            __threadfence();
        }
        grid.sync();
    }
    /* keys above the prefix are selected, keys that match it are the candidates for the rest */
    {
        const int bits_done = vs->bits_done;
        const cv_u128 prefix = ((cv_u128)vs->prefix_hi << 64) | vs->prefix_lo;
        for (long long i = g0; i < n; i += gstride) {
            const cv_u128 key = cv_sel_key(vals[i], (unsigned int)i);
            if (bits_done == 0 || (key >> (96 - bits_done)) >= prefix) {
                const unsigned int at = atomicAdd(&st->n_cand, 1u);
                if (at < CV_SEL_KMAX + CV_SEL_CAP) {
                    cand_key[at] = (unsigned long long)(key >> 32);
                    cand_idx[at] = (unsigned int)i;
                }
            }
        }
    }
    grid.sync();
    if (blockIdx.x != 0)
        return;
    /* CTA 0: the candidates sorted by (key descending, index ascending), the first K written out */
    unsigned long long *key = reinterpret_cast<unsigned long long *>(cv_sel_raw);
    unsigned int *idx = reinterpret_cast<unsigned int *>(key + 4096);
    const int nc = (int)min((unsigned int)vs->n_cand, (unsigned int)(CV_SEL_KMAX + CV_SEL_CAP));
    int m = 1;
    while (m < nc)
        m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        key[i] = i < nc ? __ldcg(cand_key + i) : 0ULL;
        idx[i] = i < nc ? __ldcg(cand_idx + i) : 0xffffffffu;
    }
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const unsigned long long ka = key[i], kb = key[j];
                    const unsigned int ia = idx[i], ib = idx[j];
                    const bool b_first = kb > ka || (kb == ka && ib < ia);
                    if (b_first == up) {
                        key[i] = kb;
                        key[j] = ka;
                        idx[i] = ib;
                        idx[j] = ia;
                    }
                }
            }
            __syncthreads();
        }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        if (k >= nc || k >= n) {
            out_v[k] = -INFINITY;
            out_i[k] = -1;
        } else {
            const unsigned long long kk = key[k];
            const unsigned long long b = (kk >> 63) ? (kk & 0x7fffffffffffffffULL) : ~kk;
            out_v[k] = __longlong_as_double((long long)b);
            out_i[k] = (long long)idx[k];
        }
    }
}

size_t cv_topk_select_bytes(void)
{
    return 256 + CV_SEL_BINS * sizeof(unsigned int) + (size_t)(CV_SEL_KMAX + CV_SEL_CAP) * 12 + 256;
}

bool cv_topk_select_fits(long long n, int K) { return K <= CV_SEL_KMAX && n <= 0xffffffffLL && n > 0; }

cudaError_t cv_launch_topk_radix_select(const double *ll, long long n, int K, void *scratch, int n_sm,
                                        double *out_ll, long long *out_idx, cudaStream_t stream)
{
    unsigned char *base = (unsigned char *)scratch;
    CvSelState *st = (CvSelState *)base;
    unsigned int *hist = (unsigned int *)(base + 256);
    unsigned long long *cand_key = (unsigned long long *)(base + 256 + CV_SEL_BINS * sizeof(unsigned int));
    unsigned int *cand_idx = (unsigned int *)(cand_key + CV_SEL_KMAX + CV_SEL_CAP);
    const size_t smem = 4096 * 12; /* the histogram and suffix sums of a round, later the candidates */
    static int per_sm = -1;
    if (per_sm < 0) {
        cudaError_t e = cudaFuncSetAttribute(cv_sel_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cv_sel_fused, 512, smem);
        if (e != cudaSuccess)
            return e;
        per_sm = occ < 1 ? 1 : (occ > 2 ? 2 : occ);
    }
    long long blocks = (n + 511) / 512;
    if (blocks > (long long)per_sm * n_sm)
        blocks = (long long)per_sm * n_sm; /* all CTAs resident: the kernel uses grid-wide barriers */
    void *args[] = {(void *)&ll, (void *)&n, (void *)&K, (void *)&st, (void *)&hist, (void *)&cand_key,
                    (void *)&cand_idx, (void *)&out_ll, (void *)&out_idx};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)cv_sel_fused, dim3((unsigned int)blocks), dim3(512), args,
                                                smem, stream);
    if (e != cudaSuccess)
        return e;
    return cudaGetLastError();
}
