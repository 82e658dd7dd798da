/*
 * cvtables.h -- host-side construction of the histogram tables a context keeps in HBM
 * (CvTables, cvpoint.h).  Host only; used by the C-ABI library at ctx_create and by the
 * test-only emulation.
 *
 * The bins of `hist` (models.py:27; any set of non-negative keys, in the caller's dict order) are
 * covered by rows of 16 consecutive bins: the smallest uncovered key opens a row.  Slots of a row
 * whose bin is not a key of hist are padding (slot_mult = 0).  Rows are grouped in blocks of 64;
 * inside a block, runs of rows that follow each other without a gap form segments of at most
 * CV_SEGMAX rows -- the unit that is seeded by one exp() per mixture term.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "cvpoint.h"

struct CvHostTables {
    std::vector<double> row_j0, row_head_h, row_head_l, row_up, row_dn, slot_mult, slot_h;
    std::vector<double> copy_log_h, copy_log_l;
    std::vector<int> slot_bin, seg_first, seg_len, blk_seg_begin;
    int n_rows = 0, n_blocks = 0, max_bin = 0;
};

/* log(j!) as a double-double for every j that opens a row */
static inline void cv_log_factorials(const std::vector<long> &at, std::vector<cv_dd> &out)
{
    out.resize(at.size());
    cv_dd acc = {0.0, 0.0};
    long j = 1; /* acc = log(j!) */
    for (size_t i = 0; i < at.size(); i++) { /* `at` ascending */
        long want = at[i] < 1 ? 1 : at[i];
        while (j < want) {
            j++;
            acc = cv_dd_add(acc, cv_log_dd((double)j));
        }
        out[i] = acc;
    }
}

/* returns "" or an error message */
static inline std::string cv_build_tables(int n_bins, const int *bin_j, const double *bin_h,
                                          CvHostTables &T)
{
    if (n_bins <= 0)
        return "empty histogram";
    std::vector<std::pair<long, int>> keys(n_bins);
    for (int b = 0; b < n_bins; b++) {
        if (bin_j[b] < 0)
            return "negative histogram key";
        if (bin_j[b] > (1 << 24))
            return "histogram key above 2^24";
        keys[b] = std::make_pair((long)bin_j[b], b);
    }
    std::sort(keys.begin(), keys.end());
    for (int b = 1; b < n_bins; b++)
        if (keys[b].first == keys[b - 1].first)
            return "duplicate histogram key";
    T.max_bin = (int)keys.back().first;

    std::vector<long> heads;
    for (int b = 0; b < n_bins;) {
        long j0 = keys[b].first;
        heads.push_back(j0);
        while (b < n_bins && keys[b].first < j0 + CV_W)
            b++;
    }
    T.n_rows = (int)heads.size();
    T.n_blocks = (T.n_rows + CV_RB - 1) / CV_RB;
    int padded = T.n_blocks * CV_RB;
    T.row_j0.assign(padded, 0.0);
    T.row_head_h.assign(padded, 0.0);
    T.row_head_l.assign(padded, 0.0);
    T.row_up.assign(padded, 0.0);
    T.row_dn.assign(padded, 0.0);
    T.slot_mult.assign((size_t)padded * CV_W, 0.0);
    T.slot_h.assign((size_t)padded * CV_W, 0.0);
    T.slot_bin.assign((size_t)padded * CV_W, -1);

    std::vector<cv_dd> logfact;
    cv_log_factorials(heads, logfact);
    const long double unscale = expl(-(long double)CV_SCALE_LOG);
    for (int r = 0; r < T.n_rows; r++) {
        long j0 = heads[r];
        T.row_j0[r] = (double)j0;
        cv_dd head = cv_dd_add_d({-logfact[r].hi, -logfact[r].lo}, CV_SCALE_LOG);
        T.row_head_h[r] = head.hi;
        T.row_head_l[r] = head.lo;
        long double run = 1.0L; /* (j0+i)! / j0! */
        for (int i = 1; i <= CV_W; i++)
            run *= (long double)(j0 + i);
        if (r + 1 < T.n_rows && heads[r + 1] == j0 + CV_W) {
            T.row_up[r + 1] = (double)(1.0L / run);
            T.row_dn[r] = (double)run;
        }
    }
    /* slots */
    {
        int r = 0;
        for (int b = 0; b < n_bins; b++) {
            long j = keys[b].first;
            while (j >= heads[r] + CV_W)
                r++;
            int i = (int)(j - heads[r]);
            long double run = 1.0L;
            for (int m = 1; m <= i; m++)
                run *= (long double)(heads[r] + m);
            size_t slot = (size_t)r * CV_W + i;
            T.slot_mult[slot] = (double)(unscale / run);
            T.slot_h[slot] = bin_h ? bin_h[keys[b].second] : 0.0;
            T.slot_bin[slot] = keys[b].second;
        }
    }
    /* log(o) for every copy number the cut-off can reach (models.py:186: o < max(hist)) */
    T.copy_log_h.assign((size_t)T.max_bin + 2, 0.0);
    T.copy_log_l.assign((size_t)T.max_bin + 2, 0.0);
    for (int o = 2; o <= T.max_bin + 1; o++) {
        cv_dd lg = cv_log_dd((double)o);
        T.copy_log_h[o] = lg.hi;
        T.copy_log_l[o] = lg.lo;
    }
    /* segments */
    T.blk_seg_begin.assign(T.n_blocks + 1, 0);
    for (int blk = 0; blk < T.n_blocks; blk++) {
        T.blk_seg_begin[blk] = (int)T.seg_first.size();
        int lo = blk * CV_RB, hi = std::min(T.n_rows, lo + CV_RB);
        int r = lo;
        while (r < hi) {
            int start = r;
            r++;
            while (r < hi && r - start < CV_SEGMAX && heads[r] == heads[r - 1] + CV_W)
                r++;
            T.seg_first.push_back(start - lo);
            T.seg_len.push_back(r - start);
        }
    }
    T.blk_seg_begin[T.n_blocks] = (int)T.seg_first.size();
    return "";
}

static inline CvTables cv_tables_view(const CvHostTables &T)
{
    CvTables v;
    v.row_j0 = T.row_j0.data();
    v.row_head_h = T.row_head_h.data();
    v.row_head_l = T.row_head_l.data();
    v.row_up = T.row_up.data();
    v.row_dn = T.row_dn.data();
    v.slot_mult = T.slot_mult.data();
    v.slot_h = T.slot_h.data();
    v.slot_bin = T.slot_bin.data();
    v.copy_log_h = T.copy_log_h.data();
    v.copy_log_l = T.copy_log_l.data();
    v.seg_first = T.seg_first.data();
    v.seg_len = T.seg_len.data();
    v.blk_seg_begin = T.blk_seg_begin.data();
    return v;
}
