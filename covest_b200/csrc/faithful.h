/*
 * faithful.h -- launch interface of the term-by-term re-evaluation of marked points (faithful.cu).
 */
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

struct CvFaithTables {
    const int *key;    /* [n] histogram keys, ascending */
    const double *cnt; /* [n] their counts */
    int n;
    double *scratch;   /* cv_faithful_warps(n_sm) x n doubles */
};

/* warps the kernel runs at most: sizes the scratch */
int cv_faithful_warps(int n_sm);

/* Re-evaluates every point whose value in out_ll is marked (finite and below CV_BAND_LL, cvmodel.h)
 * and overwrites it; *n_fixed (device, may be NULL) counts them. */
cudaError_t cv_launch_faithful(const CvModelDesc &m, const CvLattice &lat, const double *params, long long n,
                               int clip, double *out_ll, const CvFaithTables &ft, int n_sm,
                               unsigned long long *n_fixed, cudaStream_t stream);

/* marks every point: the whole batch then goes through the term-by-term evaluation (path mode 5,
 * a check of that kernel against the oracle on ordinary points) */
cudaError_t cv_launch_mark_all(double *out_ll, long long n, cudaStream_t stream);
