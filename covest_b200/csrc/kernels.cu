/*
 * kernels.cu -- hand-written sm_100a kernels of the CovEst likelihood path.
 *
 *   cv_loglik_kernel   K1/K2: one 128-thread CTA per parameter point, four CTAs per SM (persistent
 *                      CTAs drawing point indices
 *                      from a device counter), phases of cvpoint.h separated by __syncthreads().
 *                      FP64 throughout; the inner loop is one DFMA per (mixture term, bin).
 *   cv_topk_select     K3: deterministic top-K of the log-likelihoods.
 *   cv_gather_rows     (ll, params...) rows of the selected points.
 *   cv_peak_probe      register-resident DFMA / DMMA chains: the measured FP64 roofline.
 *
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 (covest_b200/build.py)
 */
#include "kernels.h"

#define CV_FULL_MASK 0xffffffffu

__device__ __forceinline__ void cv_lattice_point(const CvLattice &lat, long long i, double *row)
{
    long long idx = lat.first + i * lat.stride;
#pragma unroll
    for (int a = CV_MAX_PARAMS - 1; a >= 0; a--) {
        if (a < lat.n_axes) {
            int n = lat.len[a];
            long long q = idx / n;
            row[a] = lat.axis[a][(int)(idx - q * n)];
            idx = q;
        }
    }
}

#define CV_CTAS_PER_SM 4

__device__ __forceinline__ CvPartial cv_partial_shfl_down(const CvPartial &p, int delta)
{
    CvPartial q;
    q.sum_h = __shfl_down_sync(CV_FULL_MASK, p.sum_h, delta);
    q.sum_l = __shfl_down_sync(CV_FULL_MASK, p.sum_l, delta);
    q.mass_h = __shfl_down_sync(CV_FULL_MASK, p.mass_h, delta);
    q.mass_l = __shfl_down_sync(CV_FULL_MASK, p.mass_l, delta);
    return q;
}

__global__ void __launch_bounds__(CV_NT, CV_CTAS_PER_SM)
cv_loglik_kernel(const __grid_constant__ CvModelDesc m, const __grid_constant__ CvLattice lat,
                 const double *__restrict__ params, long long n_points, int clip,
                 double *__restrict__ out_ll, double *__restrict__ out_probs,
                 unsigned long long *counter)
{
    extern __shared__ __align__(16) unsigned char cv_smem_raw[];
    CvPointShared &sh = *reinterpret_cast<CvPointShared *>(cv_smem_raw);
    __shared__ long long s_point;
    __shared__ double s_row[CV_MAX_PARAMS];

    const int tid = threadIdx.x;
    const int S = m.n_err;
    const int cpt = cv_copies_per_tile(S);

    for (;;) {
        if (tid == 0)
            s_point = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long point = s_point;
        if (point >= n_points)
            break;
        const double *row;
        if (lat.enabled) {
            if (tid == 0)
                cv_lattice_point(lat, point, s_row);
            __syncthreads();
            row = s_row;
        } else {
            row = params + point * m.n_param;
        }
        cv_phase_header(tid, m, row, clip, sh);
        __syncthreads();
        if (m.model_kind) { /* models.py:185-191 */
            for (int first = 1; first < m.max_bin; first += CV_NT) {
                int cand = cv_phase_cut_candidate(tid, m, sh, first);
                if (cand != 0x7fffffff)
                    atomicMin(&sh.o_end, cand);
                __syncthreads();
                int o_end_now = sh.o_end;
                __syncthreads();
                if (o_end_now < first + CV_NT)
                    break;
            }
        }
        const int o_end = sh.o_end;
        const bool single_tile = (o_end - 1) <= cpt;
        CvPartial part = {0.0, 0.0, 0.0, 0.0};
        double *probs_row = out_probs ? out_probs + point * (long long)m.n_bins : nullptr;

        for (int blk = 0; blk < m.n_blocks; blk++) {
            const int nrows_blk = min(CV_RB, m.n_rows - blk * CV_RB);
            double acc[32];
#pragma unroll
            for (int i = 0; i < 32; i++)
                acc[i] = 0.0;
            for (int tile_o = 1; tile_o < o_end; tile_o += cpt) {
                const int ncop = min(cpt, o_end - tile_o);
                const int nterms = ncop * S;
                if (!(single_tile && blk > 0)) { /* the term constants of a lone tile are kept */
                    cv_phase_mass(tid, m, tile_o, nterms, sh);
                    __syncthreads();
                    cv_phase_terms(tid, m, tile_o, nterms, sh);
                    __syncthreads();
                    cv_phase_powers(tid, CV_NT, nterms, sh);
                }
                cv_phase_seeds(tid, CV_NT, m, blk, nterms, sh);
                __syncthreads();
                cv_phase_fma(tid, nterms, nrows_blk, sh, acc);
                __syncthreads();
            }
            cv_phase_spill(tid, sh, acc);
            __syncthreads();
            cv_phase_epilogue(tid, m, blk, sh, part, probs_row);
            __syncthreads();
        }
        /* block reduction of the partial sums, in a fixed order: warp 0 folds the published
         * partials, then a shuffle tree */
        cv_phase_publish(tid, sh, part);
        __syncthreads();
        if (tid < 32) {
            CvPartial total = cv_phase_fold(tid, sh);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                CvPartial q = cv_partial_shfl_down(total, d);
                cv_partial_merge(total, q);
            }
            if (tid == 0)
                out_ll[point] = cv_point_finish(m, total);
        }
    }
}

int cv_loglik_smem_bytes() { return (int)sizeof(CvPointShared); }

cudaError_t cv_launch_loglik(const CvModelDesc &m, const CvLattice &lat, const double *params,
                             long long n_points, int clip, double *out_ll, double *out_probs,
                             unsigned long long *counter, int n_sm, cudaStream_t stream)
{
    /* per device, and cheap: set on every launch */
    cudaError_t e = cudaFuncSetAttribute(cv_loglik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         cv_loglik_smem_bytes());
    if (e != cudaSuccess)
        return e;
    if (n_points <= 0)
        return cudaSuccess;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess)
        return e;
    long long want = (long long)CV_CTAS_PER_SM * n_sm; /* resident CTAs */
    int grid = (int)(n_points < want ? n_points : want);
    cv_loglik_kernel<<<grid, CV_NT, cv_loglik_smem_bytes(), stream>>>(m, lat, params, n_points, clip,
                                                                      out_ll, out_probs, counter);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------- */
/* K3: top-K.  Total order: larger ll first, then lower index; NaN counts as -inf.              */
/* ------------------------------------------------------------------------------------------- */
struct CvKey {
    double v;
    long long i;
};

__device__ __forceinline__ bool cv_key_before(const CvKey &a, const CvKey &b)
{
    return a.v > b.v || (a.v == b.v && a.i < b.i);
}

__device__ __forceinline__ double cv_key_value(double v) { return (v != v) ? -INFINITY : v; }

/* Each CTA selects the K best keys of its slice [cta * chunk, ...) of (vals, idxs).  idxs may be
 * null (index = position).  Output: out_v/out_i[cta * K + k], descending; missing = (-inf, -1)
 * ordered last (index LLONG_MAX internally). */
__global__ void __launch_bounds__(256)
cv_topk_select(const double *__restrict__ vals, const long long *__restrict__ idxs, long long n,
               long long chunk, int K, double *__restrict__ out_v, long long *__restrict__ out_i)
{
    __shared__ CvKey s_best[8];
    __shared__ CvKey s_prev;
    const int tid = threadIdx.x;
    long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk < n ? lo + chunk : n;
    CvKey prev = {INFINITY, -1}; /* before every real key */
    for (int k = 0; k < K; k++) {
        CvKey best = {-INFINITY, 0x7fffffffffffffffLL};
        for (long long p = lo + tid; p < hi; p += blockDim.x) {
            CvKey c = {cv_key_value(vals[p]), idxs ? idxs[p] : p};
            if (c.i < 0)
                continue; /* padding of an earlier stage */
            if (cv_key_before(prev, c) && cv_key_before(c, best))
                best = c;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            CvKey o;
            o.v = __shfl_down_sync(CV_FULL_MASK, best.v, d);
            o.i = __shfl_down_sync(CV_FULL_MASK, best.i, d);
            if (cv_key_before(o, best))
                best = o;
        }
        if ((tid & 31) == 0)
            s_best[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            CvKey b = s_best[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); w++)
                if (cv_key_before(s_best[w], b))
                    b = s_best[w];
            s_prev = b;
            bool none = (b.i == 0x7fffffffffffffffLL);
            out_v[(long long)blockIdx.x * K + k] = none ? -INFINITY : b.v;
            out_i[(long long)blockIdx.x * K + k] = none ? -1 : b.i;
        }
        __syncthreads();
        prev = s_prev;
        __syncthreads();
    }
}

cudaError_t cv_launch_topk(const double *ll, long long n_points, int K, double *cand_ll,
                           long long *cand_idx, int n_cta, double *out_ll, long long *out_idx,
                           cudaStream_t stream)
{
    if (n_cta < 1)
        n_cta = 1;
    long long chunk = (n_points + n_cta - 1) / n_cta;
    if (chunk < 1)
        chunk = 1;
    cv_topk_select<<<n_cta, 256, 0, stream>>>(ll, nullptr, n_points, chunk, K, cand_ll, cand_idx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return e;
    long long n2 = (long long)n_cta * K;
    cv_topk_select<<<1, 256, 0, stream>>>(cand_ll, cand_idx, n2, n2, K, out_ll, out_idx);
    return cudaGetLastError();
}

__global__ void cv_gather_rows(const __grid_constant__ CvLattice lat, const double *__restrict__ params,
                               int n_param, const double *__restrict__ sel_ll,
                               const long long *__restrict__ sel_idx, int K,
                               double *__restrict__ out_rows)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K)
        return;
    double *o = out_rows + (long long)k * (1 + n_param);
    long long i = sel_idx[k];
    o[0] = sel_ll[k];
    if (i < 0) {
        for (int a = 0; a < n_param; a++)
            o[1 + a] = NAN;
        return;
    }
    if (lat.enabled) {
        double row[CV_MAX_PARAMS];
        cv_lattice_point(lat, i, row);
        for (int a = 0; a < n_param; a++)
            o[1 + a] = row[a];
    } else {
        for (int a = 0; a < n_param; a++)
            o[1 + a] = params[i * n_param + a];
    }
}

cudaError_t cv_launch_gather_rows(const CvLattice &lat, const double *params, int n_param,
                                  const double *sel_ll, const long long *sel_idx, int K,
                                  double *out_rows, cudaStream_t stream)
{
    cv_gather_rows<<<(K + 127) / 128, 128, 0, stream>>>(lat, params, n_param, sel_ll, sel_idx, K,
                                                        out_rows);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------- */
/* FP64 peak probes                                                                             */
/* ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256) cv_peak_probe_dfma(int iters, double *sink)
{
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * blockIdx.x;
    double x0 = 0.1, x1 = 0.2, x2 = 0.3, x3 = 0.4, x4 = 0.5, x5 = 0.6, x6 = 0.7, x7 = 0.8;
    double y0 = 0.15, y1 = 0.25, y2 = 0.35, y3 = 0.45, y4 = 0.55, y5 = 0.65, y6 = 0.75, y7 = 0.85;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = __fma_rn(x0, a, b);
            x1 = __fma_rn(x1, a, b);
            x2 = __fma_rn(x2, a, b);
            x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b);
            x5 = __fma_rn(x5, a, b);
            x6 = __fma_rn(x6, a, b);
            x7 = __fma_rn(x7, a, b);
            y0 = __fma_rn(y0, a, b);
            y1 = __fma_rn(y1, a, b);
            y2 = __fma_rn(y2, a, b);
            y3 = __fma_rn(y3, a, b);
            y4 = __fma_rn(y4, a, b);
            y5 = __fma_rn(y5, a, b);
            y6 = __fma_rn(y6, a, b);
            y7 = __fma_rn(y7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)) + ((y0 + y1) + (y2 + y3)) +
               ((y4 + y5) + (y6 + y7));
    if (s == 12345.678)
        sink[0] = s;
}

__device__ __forceinline__ void cv_dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) cv_peak_probe_dmma(int iters, double *sink)
{
    double a = 1e-3 * (threadIdx.x & 7), b = 1e-3 * (threadIdx.x & 3);
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        c[i] = 0.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
            cv_dmma_m8n8k4(c[2 * u], c[2 * u + 1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        s += c[i];
    if (s == 12345.678)
        sink[0] = s;
}

cudaError_t cv_launch_peak_probe(int kind, int n_cta, int iters, double *sink, double *flop,
                                 cudaStream_t stream)
{
    if (kind == 0) {
        cv_peak_probe_dfma<<<n_cta, 256, 0, stream>>>(iters, sink);
        *flop = 2.0 * 16.0 * 8.0 * (double)iters * 256.0 * (double)n_cta;
    } else {
        cv_peak_probe_dmma<<<n_cta, 256, 0, stream>>>(iters, sink);
        /* one m8n8k4 = 8*8*4 FMA per warp */
        *flop = 2.0 * 256.0 * 8.0 * (double)iters * 8.0 * (double)n_cta;
    }
    return cudaGetLastError();
}
