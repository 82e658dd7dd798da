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
