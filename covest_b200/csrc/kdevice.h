/*
 * kdevice.h -- small device-side helpers shared by the kernels of kernels.cu and factored.cu.
 */
#pragma once
#include "kernels.h"

#define CV_FULL_MASK 0xffffffffu

__device__ __forceinline__ void cv_lattice_point(const CvLattice &lat, long long i, double *row)
{
    long long idx = lat.first + i * lat.stride;
    if (lat.block > 1) {
        const long long run = i / lat.block;
        idx = (lat.first + run * lat.stride) * lat.block + (i - run * lat.block);
    }
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

/* shared memory of a CTA: the group records (groups_staged * CV_GD doubles), then per warp a
 * CvWarpFixed followed by its variable part */
__host__ __device__ __forceinline__ size_t cv_warp_bytes(int n_err)
{
    size_t b = sizeof(CvWarpFixed) + (size_t)cv_warp_var_doubles(n_err) * sizeof(double);
    return (b + 15) & ~(size_t)15;
}

/* ---- helpers shared by factored.cu and faithful.cu ---- */
__device__ __forceinline__ void cvf_raw_row(const CvModelDesc &m, const CvLattice &lat,
                                            const double *__restrict__ params, long long i, double *row)
{
    if (lat.enabled) {
        cv_lattice_point(lat, i, row);
    } else {
#pragma unroll
        for (int a = 0; a < CV_MAX_PARAMS; a++)
            row[a] = a < m.n_param ? params[i * m.n_param + a] : 0.0;
    }
}

__device__ __forceinline__ double cvf_clipped(const CvModelDesc &m, const double *row, int clip, int a)
{
    return clip ? cv_clip(row[a], m.lo[a], m.hi[a]) : row[a];
}

/* models.py:185-191: the first o in [1, max(hist)) with b(o) <= threshold, else max(hist).  The
 * geometric tail b(o) = many * base^(o-3) is searched from a closed-form estimate with the exact
 * predicate; anything unusual (base outside (0, 1), non-positive threshold) is scanned. */
__host__ __device__ inline int cvf_cutoff(const CvModelDesc &m, double q1, double two, double many, double base)
{
    const double thr = m.threshold;
    const int top = m.max_bin;
    if (!(thr == thr))
        return top;
    if (1 < top && q1 <= thr)
        return 1;
    if (2 < top && two <= thr)
        return 2;
    if (top <= 3)
        return top;
    if (many <= thr)
        return 3;
    if (!(many == many) || !(base == base))
        return top; /* every comparison is false */
    if (base > 0.0 && base < 1.0 && thr > 0.0 && many - many == 0.0) {
        double est = 3.0 + log(thr / many) / log(base);
        int o = est < (double)top ? (int)est - 1 : top - 1;
        if (o < 3)
            o = 3;
        while (o < top && !(cv_copy_weight(o, q1, two, many, base) <= thr))
            o++;
        while (o > 3 && cv_copy_weight(o - 1, q1, two, many, base) <= thr)
            o--;
        return o < top ? o : top;
    }
    for (int o = 4; o < top; o++)
        if (cv_copy_weight(o, q1, two, many, base) <= thr)
            return o;
    return top;
}

