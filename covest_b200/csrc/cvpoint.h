/*
 * cvpoint.h -- the per-point evaluation of the CovEst mixture likelihood, written as a sequence
 * of *phases*.  Inside the sm_100a kernel (loglik_kernel.cu) every thread of a CTA calls each
 * phase with its own `tid` and the phases are separated by __syncthreads(); the test-only host
 * emulation (tests/host_math/emulate.cpp) calls the same functions in a serial loop over `tid`,
 * phase by phase, which is the same computation.  No phase uses warp intrinsics.
 *
 * One CTA evaluates one parameter point at a time:
 *
 *   header      clip the parameters (models.py:60-69), error-class rates l_s (models.py:71-79),
 *               copy-number weights (models.py:193-208) and the cut-off O_thr (models.py:185-191)
 *   per tile of up to CV_TT mixture terms (o, s):
 *     mass      n_os = comb[s] * (1.0 - exp(-o*l_s))                      (models.py:87, :221)
 *     terms     a_os = n_os / sum_s n_os, w = b(o) * a_os, log-domain constants of the term
 *     powers    PW[t][i] = lam_t^i, i = 0..15
 *     seeds     SD[t][row] = scaled value of term t at the head bin of every 16-bin row of the
 *               current block of rows: one exp() at the row holding the mode of the term, then
 *               the reference's own product recurrence (covest_poissonmodule.c:22-24) walked
 *               outwards 16 bins at a time
 *     fma       ACC[row][i] += SD[t][row] * PW[t][i]     -- one FP64 FMA per (term, bin)
 *   spill       per-warp partial ACCs to shared memory
 *   epilogue    p_j = ACC * slot_mult, mass += p_j, sum += h_j * log p_j  (models.py:100-107)
 *
 * Reference lines are relative to /root/reference.
 */
#pragma once
#include "cvmodel.h"

#define CV_W 16        /* bins per row (chain) */
#define CV_RB 64       /* rows per block: 1024 bins */
#define CV_TT 64       /* mixture terms per tile */
#define CV_SSTRIDE 65  /* padded row stride of the seed matrix (bank-conflict-free columns) */
#define CV_NT 128      /* threads per CTA; four CTAs share an SM */
#define CV_NWARP 4
#define CV_SEGMAX 32   /* longest run of rows seeded from one exp() */
#define CV_MAX_PARAMS 5
#define CV_BW_CACHE 512 /* copy weights b(o), o < 512, are kept from the cut-off search */

/* Histogram-side tables of a context, all indexed by row / slot (slot = row * 16 + i).  Device
 * memory in the product, host memory in the emulation.  Built by cv_build_tables (cvtables.h). */
struct CvTables {
    const double *row_j0;      /* head bin of the row */
    const double *row_head_h;  /* CV_SCALE_LOG - lgamma(j0 + 1), double-double */
    const double *row_head_l;
    const double *row_up;      /* j0[r-1]! / j0[r]!   (row r continues row r-1) */
    const double *row_dn;      /* j0[r+1]! / j0[r]!   (row r+1 continues row r) */
    const double *slot_mult;   /* exp(-CV_SCALE_LOG) * j0! / (j0+i)!; 0 marks a slot not in hist */
    const double *slot_h;      /* count h_j */
    const int *slot_bin;       /* position of the bin in the caller's hist order, -1 = padding */
    const double *copy_log_h;  /* log(o), o = 0..max_bin, as a double-double (entry 0 unused) */
    const double *copy_log_l;
    const int *seg_first;      /* segments: runs of consecutive rows inside one block */
    const int *seg_len;
    const int *blk_seg_begin;  /* [n_blocks + 1] */
};

struct CvModelDesc {
    int model_kind; /* 0 basic (models.py:17), 1 repeats (models.py:173) */
    int k, r;
    int n_err;      /* max_error (models.py:28-31) */
    int n_param;
    int n_bins, n_rows, n_blocks;
    int max_bin;    /* max(hist), models.py:186 */
    double tail;
    double threshold; /* NaN = None */
    double lo[CV_MAX_PARAMS], hi[CV_MAX_PARAMS]; /* NaN = open */
    double comb[CV_MAX_ERR];  /* models.py:25, as the host computed it */
    double pow3[CV_MAX_ERR];  /* 3 ** -s as the host computed it */
    CvTables tab;
};

/* Working set of one point; shared memory on the device. */
struct CvPointShared {
    double par[CV_MAX_PARAMS];
    double ls[CV_MAX_ERR];                 /* l_s */
    double lls_h[CV_MAX_ERR], lls_l[CV_MAX_ERR]; /* log(l_s), double-double */
    double two, many, base; /* (1-q1)*q2, (1-q1)*(1-q2)*q, 1-q */
    int o_end;              /* O_thr: copies 1 .. o_end-1 are evaluated */
    int pad_;
    double bw[CV_BW_CACHE]; /* b(o) */
    double nmass[CV_TT];
    double lam[CV_TT], lh[CV_TT], ll[CV_TT], lin[CV_TT], f[CV_TT];
    double l2[CV_TT], l4[CV_TT], l8[CV_TT], pw16[CV_TT], ipw16[CV_TT];
    double PW[CV_TT * CV_W];
    double SD[CV_TT * CV_SSTRIDE]; /* seeds; reused as the cross-warp reduction buffer and, after
                                      the last epilogue, for the final reduction */
};

struct CvPartial {
    double sum_h, sum_l;   /* sum_j h_j log p_j, compensated */
    double mass_h, mass_l; /* sum_j p_j, compensated */
};

/* copies per tile */
CV_HD int cv_copies_per_tile(int n_err) { return CV_TT / n_err; }

/* ---- header ------------------------------------------------------------------------------- */
CV_HD void cv_phase_header(int tid, const CvModelDesc &m, const double *row, int clip,
                           CvPointShared &sh)
{
    if (tid < m.n_err) {
        double c = row[0], e = row[1];
        if (clip) {
            c = cv_clip(c, m.lo[0], m.hi[0]);
            e = cv_clip(e, m.lo[1], m.hi[1]);
        }
        double ck = cv_kmer_coverage(c, m.k, m.r);
        double l = cv_error_class_rate(ck, m.pow3[tid], e, m.k, tid);
        sh.ls[tid] = l;
        cv_dd lg = {0.0, 0.0};
        if (l > 0.0 && l - l == 0.0) /* positive and finite */
            lg = cv_log_dd(l);
        sh.lls_h[tid] = lg.hi;
        sh.lls_l[tid] = lg.lo;
    }
    if (tid == CV_NT - 1) {
        for (int i = 0; i < m.n_param; i++)
            sh.par[i] = clip ? cv_clip(row[i], m.lo[i], m.hi[i]) : row[i];
        if (m.model_kind) {
            double q1 = sh.par[2], q2 = sh.par[3], q = sh.par[4];
            sh.two = cv_mul(cv_sub(1.0, q1), q2);                          /* models.py:195 */
            sh.many = cv_mul(cv_mul(cv_sub(1.0, q1), cv_sub(1.0, q2)), q); /* models.py:196 */
            sh.base = cv_sub(1.0, q);
            sh.o_end = m.max_bin; /* models.py:191 */
        } else {
            sh.two = sh.many = sh.base = 0.0;
            sh.o_end = 2; /* the basic model is the single copy o = 1 with weight 1 */
        }
    }
}

/* b(o), models.py:198-206 */
CV_HD double cv_point_copy_weight(const CvModelDesc &m, const CvPointShared &sh, int o)
{
    if (!m.model_kind)
        return 1.0;
    return cv_copy_weight(o, sh.par[2], sh.two, sh.many, sh.base);
}

/* One pass of the cut-off search models.py:187-190 over copies first_o .. first_o + CV_NT - 1.
 * Returns the candidate this thread found (or INT_MAX); the caller min-reduces into sh.o_end. */
CV_HD int cv_phase_cut_candidate(int tid, const CvModelDesc &m, CvPointShared &sh, int first_o)
{
    int o = first_o + tid;
    if (o >= m.max_bin)
        return 0x7fffffff;
    double b = cv_point_copy_weight(m, sh, o);
    if (o < CV_BW_CACHE)
        sh.bw[o] = b;
    if (b <= m.threshold)
        return o;
    return 0x7fffffff;
}

/* ---- per tile ----------------------------------------------------------------------------- */
/* models.py:87 / :221.  Term t of the tile is copy o = tile_o + t / S, error class s = t % S. */
CV_HD void cv_phase_mass(int tid, const CvModelDesc &m, int tile_o, int nterms, CvPointShared &sh)
{
    if (tid >= nterms)
        return;
    int S = m.n_err;
    int o = tile_o + tid / S, s = tid % S;
    double lam = cv_mul((double)o, sh.ls[s]); /* o * l_s, models.py:238 */
    sh.lam[tid] = lam;
    sh.nmass[tid] = cv_class_mass(m.comb[s], lam);
}

CV_HD void cv_phase_terms(int tid, const CvModelDesc &m, int tile_o, int nterms, CvPointShared &sh)
{
    if (tid >= nterms)
        return;
    int S = m.n_err;
    int g = tid / S, s = tid - g * S;
    int o = tile_o + g;
    /* models.py:88 / :224: Python sum(), left to right starting from int 0 */
    double total = 0.0;
    for (int i = 0; i < S; i++)
        total = cv_add(total, sh.nmass[g * S + i]);
    if (total == 0.0)
        total = 1.0; /* utils.py:25-29 fix_zero */
    double b = 1.0;
    if (m.model_kind)
        b = (o < CV_BW_CACHE) ? sh.bw[o] : cv_point_copy_weight(m, sh, o);
    double lam = sh.lam[tid];
    /* log(lam) = log(o) + log(l_s) + log(lam / (o*l_s)); the last part undoes the rounding of the
     * product lam = RN(o*l_s) and is -(o*l_s - lam)/lam to first order (|.| <= 2^-53) */
    double ls = sh.ls[s];
    double resid = cv_fma((double)o, ls, -lam);
    cv_dd lo_ = {m.tab.copy_log_h[o], m.tab.copy_log_l[o]};
    cv_dd ll_ = {sh.lls_h[s], sh.lls_l[s]};
    cv_dd lg = cv_dd_add(lo_, ll_);
    if (resid != 0.0)
        lg = cv_dd_add_d(lg, -cv_div(resid, lam));
    CvTerm t = cv_term_make(lam, cv_mul(b, sh.nmass[tid]), total, lg.hi, lg.lo);
    sh.lam[tid] = t.lam;
    sh.lh[tid] = t.lh;
    sh.ll[tid] = t.ll;
    sh.lin[tid] = t.lin;
    sh.f[tid] = t.f;
    double l2 = cv_mul(t.lam, t.lam);
    double l4 = cv_mul(l2, l2);
    double l8 = cv_mul(l4, l4);
    double l16 = cv_mul(l8, l8);
    sh.l2[tid] = l2;
    sh.l4[tid] = l4;
    sh.l8[tid] = l8;
    sh.pw16[tid] = l16;
    sh.ipw16[tid] = cv_div(1.0, l16);
}

/* PW[t][i] = lam_t^i, i < 16, from the squarings: every element is at most three products */
CV_HD void cv_phase_powers(int tid, int nthreads, int nterms, CvPointShared &sh)
{
    for (int e = tid; e < nterms * CV_W; e += nthreads) {
        int t = e >> 4, i = e & 15;
        double v = (i & 1) ? sh.lam[t] : 1.0;
        if (i & 2)
            v = cv_mul(v, sh.l2[t]);
        if (i & 4)
            v = cv_mul(v, sh.l4[t]);
        if (i & 8)
            v = cv_mul(v, sh.l8[t]);
        sh.PW[e] = v;
    }
}

/* Seeds of every (term, row) of block `blk`.  A work item is (term, segment); it takes one exp()
 * at the row of the segment that holds the mode of the term (j ~ lam) and walks outwards, where
 * the term only decreases, so that a value that underflowed never has to grow back. */
CV_HD void cv_phase_seeds(int tid, int nthreads, const CvModelDesc &m, int blk, int nterms,
                          CvPointShared &sh)
{
    const CvTables &T = m.tab;
    int sb = T.blk_seg_begin[blk];
    int nseg = T.blk_seg_begin[blk + 1] - sb;
    int items = nterms * nseg;
    for (int it = tid; it < items; it += nthreads) {
        int t = it % nterms;
        int sg = sb + it / nterms;
        int first = T.seg_first[sg], len = T.seg_len[sg];
        int grow = blk * CV_RB + first;
        double lam = sh.lam[t];
        double off = (lam - T.row_j0[grow]) * (1.0 / CV_W);
        int rs = 0;
        if (off >= (double)(len - 1))
            rs = len - 1;
        else if (off > 0.0)
            rs = (int)off;
        double seed = cv_seed(T.row_j0[grow + rs], T.row_head_h[grow + rs], T.row_head_l[grow + rs],
                              sh.lh[t], sh.ll[t], sh.lin[t], sh.f[t]);
        double *srow = sh.SD + t * CV_SSTRIDE + first;
        srow[rs] = seed;
        double v = seed;
        double step = sh.pw16[t];
        for (int r = rs + 1; r < len; r++) {
            v = cv_mul(cv_mul(v, step), T.row_up[grow + r]);
            srow[r] = v;
        }
        v = seed;
        step = sh.ipw16[t];
        for (int r = rs - 1; r >= 0; r--) {
            v = cv_mul(cv_mul(v, step), T.row_dn[grow + r]);
            srow[r] = v;
        }
    }
}

/* The accumulators of a thread: rows rg + 16a (a < 4), columns 8cg + b (b < 8) of the block, with
 * rg = lane >> 1, cg = lane & 1; warp w takes the terms t = w, w + CV_NWARP, ... (split over terms,
 * summed in cv_phase_epilogue). */
CV_HD void cv_phase_fma(int tid, int nterms, int nrows_blk, const CvPointShared &sh, double *acc)
{
    int warp = tid >> 5, lane = tid & 31;
    int rg = lane >> 1, cg = lane & 1;
    bool live0 = rg < nrows_blk, live1 = rg + 16 < nrows_blk, live2 = rg + 32 < nrows_blk,
         live3 = rg + 48 < nrows_blk;
    for (int t = warp; t < nterms; t += CV_NWARP) {
        const double *srow = sh.SD + t * CV_SSTRIDE + rg;
        const double *pw = sh.PW + t * CV_W + cg * 8;
        double b[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            b[i] = pw[i];
        double a0 = live0 ? srow[0] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++)
            acc[i] = cv_fma(a0, b[i], acc[i]);
        if (nrows_blk > 16) {
            double a1 = live1 ? srow[16] : 0.0;
            double a2 = live2 ? srow[32] : 0.0;
            double a3 = live3 ? srow[48] : 0.0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                acc[8 + i] = cv_fma(a1, b[i], acc[8 + i]);
                acc[16 + i] = cv_fma(a2, b[i], acc[16 + i]);
                acc[24 + i] = cv_fma(a3, b[i], acc[24 + i]);
            }
        }
    }
}

/* reduction buffer: red[warp][a][b][lane] */
CV_HD void cv_phase_spill(int tid, CvPointShared &sh, const double *acc)
{
    int warp = tid >> 5, lane = tid & 31;
    double *red = sh.SD + warp * (CV_RB * CV_W);
#pragma unroll
    for (int e = 0; e < 32; e++)
        red[e * 32 + lane] = acc[e];
}

CV_HD void cv_partial_add_mass(CvPartial &p, double x)
{
    cv_dd s = cv_two_sum(p.mass_h, x);
    p.mass_h = s.hi;
    p.mass_l = cv_add(p.mass_l, s.lo);
}

CV_HD void cv_partial_add_sum(CvPartial &p, double x)
{
    cv_dd s = cv_two_sum(p.sum_h, x);
    p.sum_h = s.hi;
    if (s.lo == s.lo) /* an infinite term makes the error term NaN; the sum itself stays right */
        p.sum_l = cv_add(p.sum_l, s.lo);
}

CV_HD void cv_partial_merge(CvPartial &p, const CvPartial &q)
{
    cv_partial_add_mass(p, q.mass_h);
    p.mass_l = cv_add(p.mass_l, q.mass_l);
    cv_partial_add_sum(p, q.sum_h);
    p.sum_l = cv_add(p.sum_l, q.sum_l);
}

/* models.py:100-107 per bin.  Thread tid finishes the slots e = tid + CV_NT n of the [a][b][lane]
 * order of cv_phase_spill. */
CV_HD void cv_phase_epilogue(int tid, const CvModelDesc &m, int blk, const CvPointShared &sh,
                             CvPartial &part, double *out_probs)
{
    const CvTables &T = m.tab;
    for (int n = 0; n < (CV_RB * CV_W) / CV_NT; n++) {
        int e = tid + n * CV_NT;
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < CV_NWARP; w++)
            v = cv_add(v, sh.SD[w * (CV_RB * CV_W) + e]);
        int lane = e & 31, b = (e >> 5) & 7, a = e >> 8;
        int row = (lane >> 1) + 16 * a, col = 8 * (lane & 1) + b;
        int slot = (blk * CV_RB + row) * CV_W + col;
        double mult = T.slot_mult[slot];
        if (mult == 0.0)
            continue; /* a bin that is not in hist */
        double p = cv_mul(v, mult);
        if (out_probs)
            out_probs[T.slot_bin[slot]] = p;
        cv_partial_add_mass(part, p);
        double h = T.slot_h[slot];
        if (h != 0.0) { /* models.py:106 `if h` */
            double lg = (p <= 0.0) ? -INFINITY : log(p); /* utils.py:32-35 safe_log */
            cv_partial_add_sum(part, cv_mul(h, lg));
        }
    }
}

/* models.py:103-107 */
CV_HD double cv_point_finish(const CvModelDesc &m, const CvPartial &part)
{
    double mass = cv_add(part.mass_h, part.mass_l);
    if (!(mass < 1.0))
        mass = 1.0; /* min(1, fsum(...)): keeps the 1 unless the sum is smaller (also for NaN) */
    double sum = part.sum_h;
    if (part.sum_h - part.sum_h == 0.0) /* finite */
        sum = cv_add(part.sum_h, part.sum_l);
    return cv_finish_loglik(sum, mass, m.tail);
}

/* Final reduction, in a fixed order: every thread publishes its partial, warp 0 folds
 * CV_NT / 32 partials per lane; the caller finishes with 32 values (lane order). */
CV_HD void cv_phase_publish(int tid, CvPointShared &sh, const CvPartial &part)
{
    double *red = sh.SD; /* free after the last epilogue (the caller synchronises) */
    red[tid] = part.sum_h;
    red[CV_NT + tid] = part.sum_l;
    red[2 * CV_NT + tid] = part.mass_h;
    red[3 * CV_NT + tid] = part.mass_l;
}

CV_HD CvPartial cv_phase_fold(int lane, const CvPointShared &sh)
{
    const double *red = sh.SD;
    CvPartial acc = {red[lane], red[CV_NT + lane], red[2 * CV_NT + lane], red[3 * CV_NT + lane]};
    for (int w = 1; w < CV_NWARP; w++) {
        int i = lane + 32 * w;
        CvPartial q = {red[i], red[CV_NT + i], red[2 * CV_NT + i], red[3 * CV_NT + i]};
        cv_partial_merge(acc, q);
    }
    return acc;
}
