/*
 * faithful.h -- launch interface of the term-by-term re-evaluation of marked points (faithful.cu).
 */
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

struct CvFaithTables {
    const double *cnt_of_j; /* [j_all] count of bin j + 1; < 0: j + 1 is not a key of hist */
    const double *rcp;      /* [j_all] 1 / (j + 1) */
    int j_all;              /* max(hist) */
    int j_counted;          /* the largest key with a non-zero count */
    double *scratch;        /* cv_faithful_warps(n_sm) x acc_doubles */
    long long acc_doubles;  /* per warp: j_all rounded up to a multiple of 32 */
};

/* warps the kernel runs at most: sizes the scratch */
int cv_faithful_warps(int n_sm);

/* Re-evaluates every point whose value in out_ll is marked (finite and below CV_BAND_LL, cvmodel.h)
 * and overwrites it.  list: 2 n entries of scratch (marked points, those of them with many copy
 * numbers); counters: four device words (marked points of the call, cursor, points with many copies,
 * cursor), zeroed here. */
cudaError_t cv_launch_faithful(const CvModelDesc &m, const CvLattice &lat, const double *params, long long n,
                               int clip, double *out_ll, const CvFaithTables &ft, int n_sm, unsigned int *list,
                               unsigned long long *counters, cudaStream_t stream);

/* marks every point: the whole batch then goes through the term-by-term evaluation (path mode 5,
 * a check of that kernel against the oracle on ordinary points) */
cudaError_t cv_launch_mark_all(double *out_ll, long long n, cudaStream_t stream);
