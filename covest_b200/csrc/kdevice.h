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

