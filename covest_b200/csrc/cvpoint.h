/*
 * cvpoint.h -- the per-point evaluation of the CovEst mixture likelihood, written as a sequence
 * of warp-wide *phases*.  Inside the sm_100a kernel (kernels.cu) ONE WARP evaluates one parameter
 * point at a time: its 32 lanes call each phase with their own `lane`, the phases are separated by
 * __syncwarp(), and nothing is shared between the warps of a CTA except read-only row tables.  The
 * test-only host emulation (tests/host_math/emulate.cpp) calls the same functions in a serial loop
 * over `lane`, phase by phase, which is the same computation.  No phase uses warp intrinsics.
 *
 *   header      clip the parameters (models.py:60-69), error-class rates l_s (models.py:71-79),
 *               the constants of the copy-number weights (models.py:193-208)
 *   per pass of 32 copy numbers o: the weights b(o), one per lane, and the cut-off O_thr
 *               (models.py:185-191) -- the weights stay in registers and reach the terms by shuffle
 *   per group of up to 32 mixture terms (o, s)  [64 when there are more than 32 error classes]:
 *     mass      n_os = comb[s] * (1.0 - exp(-o*l_s))                      (models.py:87, :221)
 *     terms     a_os = n_os / sum_s n_os, w = b(o) * a_os, log-domain constants of the term
 *     per half-tile of 16 terms:
 *       powers  PW[i][t] = lam_t^i, i = 0..15
 *       seeds   SD[row][t] = scaled value of term t at the head bin of every 16-bin row of the
 *               current block of 64 rows: one exp() at the row holding the mode of the term, then
 *               the reference's own product recurrence (covest_poissonmodule.c:22-24) walked
 *               outwards 16 bins at a time
 *       fma     ACC[row][i] += SD[row][t] * PW[i][t]   -- one FP64 FMA per (term, bin), issued as
 *               m8n8k4 FP64 tensor-core MMAs: the 64 x 16 block is 8 x 2 MMA tiles, a lane holds
 *               two accumulators of each (32 in registers)
 *   epilogue    p_j = ACC * slot_mult, mass += p_j, sum += h_j * log p_j  (models.py:100-107),
 *               straight from the accumulator registers
 *
 * Reference lines are relative to /root/reference.
 */
#pragma once
#include "cvmodel.h"

#define CV_W 16        /* bins per row (chain) */
#define CV_RB 64       /* rows per block: 1024 bins */
#define CV_HT 16       /* mixture terms per half-tile (powers / seeds / fma) */
#define CV_CT 32       /* mixture terms per constant tile (terms phase: one per lane) */
#define CV_SDS 24      /* row stride of the seed matrix SD[row][term] and of the power matrix
                          PW[column][term]: 192 B, so that the 16-byte chunks an 8-lane quarter
                          warp reads for an MMA fragment (2 rows x 4 term pairs) fall into 8
                          different bank groups */
#define CV_PWS 24
#define CV_SEGMAX 32   /* longest run of rows seeded from one exp() */
#define CV_MAX_PARAMS 5
#define CV_WARPS_MAX 12 /* warps (= points in flight) per CTA; one CTA per SM */

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

/* Fixed-size part of the working set of one warp (= one point in flight); shared memory on the
 * device.  Compile-time offsets keep the hot phases free of address arithmetic. */
struct CvWarpFixed {
    double SD[CV_RB * CV_SDS];  /* seeds of the current half-tile, [row][term] */
    double PW[CV_W * CV_PWS];   /* powers of the current half-tile, [column i][term] */
    double lam[CV_CT], lh[CV_CT], ll[CV_CT], lin[CV_CT], f[CV_CT]; /* constants of a tile of terms */
    double pw16[CV_HT], ipw16[CV_HT];                              /* lam^16 and its inverse */
    double par[CV_MAX_PARAMS];
    double two, many, base; /* (1-q1)*q2, (1-q1)*(1-q2)*q, 1-q */
};

/* The working set of a warp: the fixed part, the arrays whose length depends on the number of
 * error classes, and the row tables every warp of the CTA reads. */
struct CvWarpMem {
    CvWarpFixed *fx;
    double *ls;            /* [S] l_s */
    double *lls_h, *lls_l; /* [S] log(l_s), double-double */
    double *nmass;         /* [group terms] n_os */
    double *glam;          /* [group terms] o * l_s */
    /* the row tables of CvTables every warp of the CTA reads (staged in shared memory) */
    const double *row_j0, *row_head_h, *row_head_l, *row_up, *row_dn;
};

/* terms of the largest group: whole copies, at most 32 terms unless one copy alone has more */
CV_HD int cv_copies_per_group(int n_err) { return n_err >= CV_CT ? 1 : CV_CT / n_err; }
CV_HD int cv_group_terms_max(int n_err) { return n_err > CV_CT ? 2 * CV_CT : CV_CT; }
/* doubles of the variable part of a warp's working set */
CV_HD int cv_warp_var_doubles(int n_err) { return 3 * n_err + 2 * cv_group_terms_max(n_err); }

CV_HD void cv_warp_mem_carve(CvWarpMem &M, CvWarpFixed *fx, double *var, int n_err)
{
    int gm = cv_group_terms_max(n_err);
    M.fx = fx;
    M.ls = var;
    M.lls_h = var + n_err;
    M.lls_l = var + 2 * n_err;
    M.nmass = var + 3 * n_err;
    M.glam = M.nmass + gm;
}

struct CvPartial {
    double sum_h, sum_l;   /* sum_j h_j log p_j, compensated */
    double mass_h, mass_l; /* sum_j p_j, compensated */
};

/* ---- header ------------------------------------------------------------------------------- */
/* `row` holds the raw parameters of the point (every lane has them). */
CV_HD void cv_w_header(int lane, const CvModelDesc &m, const double *row, int clip, CvWarpMem &M)
{
    CvWarpFixed &F = *M.fx;
    double c = row[0], e = row[1];
    if (clip) {
        c = cv_clip(c, m.lo[0], m.hi[0]);
        e = cv_clip(e, m.lo[1], m.hi[1]);
    }
    double ck = cv_kmer_coverage(c, m.k, m.r);
    for (int s = lane; s < m.n_err; s += 32) {
        double l = cv_error_class_rate(ck, m.pow3[s], e, m.k, s);
        M.ls[s] = l;
        cv_dd lg = {0.0, 0.0};
        if (l > 0.0 && l - l == 0.0) /* positive and finite */
            lg = cv_log_dd(l);
        M.lls_h[s] = lg.hi;
        M.lls_l[s] = lg.lo;
    }
    if (lane == 31) {
        double par[CV_MAX_PARAMS];
#pragma unroll
        for (int i = 0; i < CV_MAX_PARAMS; i++) {
            par[i] = 0.0;
            if (i < m.n_param)
                par[i] = clip ? cv_clip(row[i], m.lo[i], m.hi[i]) : row[i];
            F.par[i] = par[i];
        }
        if (m.model_kind) {
            double q1 = par[2], q2 = par[3], q = par[4];
            F.two = cv_mul(cv_sub(1.0, q1), q2);                          /* models.py:195 */
            F.many = cv_mul(cv_mul(cv_sub(1.0, q1), cv_sub(1.0, q2)), q); /* models.py:196 */
            F.base = cv_sub(1.0, q);
        } else {
            F.two = F.many = F.base = 0.0;
        }
    }
}

/* One lane's share of a pass of the cut-off search models.py:187-190 over copies
 * first_o .. first_o + 31: lane i looks at o = first_o + i.  Returns true when the evaluated
 * copies end BEFORE o (the caller takes the lowest such lane); *b is b(o), models.py:198-206.
 * The basic model is the single copy o = 1 with weight 1. */
CV_HD bool cv_w_copy_pass(int lane, const CvModelDesc &m, const CvWarpMem &M, int first_o, double *b)
{
    int o = first_o + lane;
    if (!m.model_kind) {
        *b = 1.0;
        return o >= 2;
    }
    if (o >= m.max_bin) { /* models.py:191: no cut found, O_thr = max(hist) */
        *b = 0.0;
        return true;
    }
    const CvWarpFixed &F = *M.fx;
    double w = cv_copy_weight(o, F.par[2], F.two, F.many, F.base);
    *b = w;
    return w <= m.threshold;
}

/* ---- per group ---------------------------------------------------------------------------- */
/* models.py:87 / :221.  Term t of the group is copy o = group_o + t / S, error class s = t % S. */
CV_HD void cv_w_mass(int lane, const CvModelDesc &m, int group_o, int nterms, CvWarpMem &M)
{
    int S = m.n_err;
    for (int t = lane; t < nterms; t += 32) {
        int g = t / S, s = t - g * S;
        double lam = cv_mul((double)(group_o + g), M.ls[s]); /* o * l_s, models.py:238 */
        M.glam[t] = lam;
        M.nmass[t] = cv_class_mass(m.comb[s], lam);
    }
}

/* Constants of the term t = sub + lane of the group (a dead term when t >= nterms); `b` is the
 * weight b(o) of the copy this lane's term belongs to. */
CV_HD void cv_w_terms(int lane, const CvModelDesc &m, int group_o, int nterms, int sub, double b,
                      CvWarpMem &M)
{
    CvWarpFixed &F = *M.fx;
    int t = sub + lane;
    if (t >= nterms) { /* padding of the last tile: contributes exactly 0 */
        F.lam[lane] = 1.0;
        F.lh[lane] = 0.0;
        F.ll[lane] = 0.0;
        F.lin[lane] = 0.0;
        F.f[lane] = 0.0;
        return;
    }
    int S = m.n_err;
    int g = t / S, s = t - g * S;
    int o = group_o + g;
    /* models.py:88 / :224: Python sum(), left to right starting from int 0 */
    double total = 0.0;
    for (int i = 0; i < S; i++)
        total = cv_add(total, M.nmass[g * S + i]);
    if (total == 0.0)
        total = 1.0; /* utils.py:25-29 fix_zero */
    double lam = M.glam[t];
    /* log(lam) = log(o) + log(l_s) + log(lam / (o*l_s)); the last part undoes the rounding of the
     * product lam = RN(o*l_s) and is -(o*l_s - lam)/lam to first order (|.| <= 2^-53) */
    double ls = M.ls[s];
    double resid = cv_fma((double)o, ls, -lam);
    cv_dd lo_ = {m.tab.copy_log_h[o], m.tab.copy_log_l[o]};
    cv_dd ll_ = {M.lls_h[s], M.lls_l[s]};
    cv_dd lg = cv_dd_add(lo_, ll_);
    if (resid != 0.0)
        lg = cv_dd_add_d(lg, -cv_div(resid, lam));
    CvTerm tm = cv_term_make(lam, cv_mul(b, M.nmass[t]), total, lg.hi, lg.lo);
    F.lam[lane] = tm.lam;
    F.lh[lane] = tm.lh;
    F.ll[lane] = tm.ll;
    F.lin[lane] = tm.lin;
    F.f[lane] = tm.f;
}

/* two adjacent doubles, one 16-byte access on the device */
struct cv_pair {
    double x, y;
};
CV_HD cv_pair cv_ld2(const double *p)
{
#if defined(__CUDA_ARCH__)
    double2 v = *reinterpret_cast<const double2 *>(p);
    cv_pair r = {v.x, v.y};
    return r;
#else
    cv_pair r = {p[0], p[1]};
    return r;
#endif
}
CV_HD void cv_st2(double *p, double x, double y)
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<double2 *>(p) = make_double2(x, y);
#else
    p[0] = x;
    p[1] = y;
#endif
}

/* ---- per half-tile ------------------------------------------------------------------------ */
/* PW[i][t] = lam_t^i, i < 16: lanes 2t and 2t+1 take the lower and the upper eight; every element
 * is at most three products of the squarings (the lower eight are multiplied by an exact 1.0 so
 * that the warp does not diverge).  Also lam^16 (kept by lane 2t) and 1 / lam^16 (lane 2t+1). */
CV_HD void cv_w_powers(int lane, int half, CvWarpMem &M)
{
    CvWarpFixed &F = *M.fx;
    int t = lane >> 1, up = lane & 1;
    double lam = F.lam[half * CV_HT + t];
    double l2 = cv_mul(lam, lam);
    double l4 = cv_mul(l2, l2);
    double l8 = cv_mul(l4, l4);
    double l3 = cv_mul(l2, lam);
    double l5 = cv_mul(l4, lam);
    double l6 = cv_mul(l4, l2);
    double l7 = cv_mul(l4, l3);
    double l16 = cv_mul(l8, l8);
    double inv16 = cv_div(1.0, l16);
    double base = up ? l8 : 1.0;
    double *pw = F.PW + up * 8 * CV_PWS + t; /* column 8 up + i of term t */
    pw[0 * CV_PWS] = base;
    pw[1 * CV_PWS] = cv_mul(base, lam);
    pw[2 * CV_PWS] = cv_mul(base, l2);
    pw[3 * CV_PWS] = cv_mul(base, l3);
    pw[4 * CV_PWS] = cv_mul(base, l4);
    pw[5 * CV_PWS] = cv_mul(base, l5);
    pw[6 * CV_PWS] = cv_mul(base, l6);
    pw[7 * CV_PWS] = cv_mul(base, l7);
    if (up)
        F.ipw16[t] = inv16;
    else
        F.pw16[t] = l16;
}

/* Seeds of every (term, row) of block `blk`.  A work item is (term, segment): one exp() at one row
 * of the segment, then the product recurrence walked over the other rows, one multiplication deep
 * per row because the factor lam^+-16 * (factorial ratio) does not depend on the running value.
 * Where the walk starts:
 *   - mode of the term (j ~ lam) at or above the top row: at the top row, walking down;
 *   - otherwise at the bottom row, walking up through the mode, when the value there is well inside
 *     the double range (it always is unless the weight of the term is minute);
 *   - otherwise at the row holding the mode, walking up and then down, so that a value that
 *     underflowed never has to grow back.
 * The first two cases share one loop with a per-lane direction, so lanes do not diverge. */
CV_HD void cv_w_seeds(int lane, const CvModelDesc &m, int blk, int half, CvWarpMem &M)
{
    CvWarpFixed &F = *M.fx;
    const CvTables &T = m.tab;
    const int sb = T.blk_seg_begin[blk];
    const int items = (T.blk_seg_begin[blk + 1] - sb) * CV_HT;
    for (int it = lane; it < items; it += 32) {
        const int t = it & (CV_HT - 1), sg = sb + (it >> 4);
        const int ct = half * CV_HT + t;
        const int first = T.seg_first[sg], top = T.seg_len[sg] - 1;
        const int grow = blk * CV_RB + first;
        const double lam = F.lam[ct], lh = F.lh[ct], lin = F.lin[ct], f = F.f[ct];
        const double j0 = M.row_j0[grow];
        const double off = (lam - j0) * (1.0 / CV_W);
        /* exponent of the scaled term at the bottom row, to a few ulps */
        const double e0 = cv_sub(cv_fma(j0, lh, M.row_head_h[grow]), lin);
        int rs;
        if (off >= (double)top)
            rs = top;
        else if (e0 > -460.0 && !(f < 0x1p-200))
            rs = 0;
        else
            rs = off > 0.0 ? (int)off : 0;
        const double seed = cv_seed(M.row_j0[grow + rs], M.row_head_h[grow + rs], M.row_head_l[grow + rs],
                                    lh, F.ll[ct], lin, f);
        double *col = F.SD + first * CV_SDS + t;
        col[rs * CV_SDS] = seed;
        const bool down = rs == top;
        const double step_dn = F.ipw16[t];
        {
            const int n1 = down ? top : top - rs;
            const int dt = down ? -1 : 1;
            const double *tp = (down ? M.row_dn : M.row_up) + grow + rs + dt;
            double *sp = col + (rs + dt) * CV_SDS;
            const double step = down ? step_dn : F.pw16[t];
            double v = seed;
            int k = 0;
            /* four rows at a time: the four table reads and factor products are issued together,
             * only the running value is a dependent chain */
            for (; k + 4 <= n1; k += 4) {
                double r0 = tp[0], r1 = tp[dt], r2 = tp[2 * dt], r3 = tp[3 * dt];
                double c0 = cv_mul(step, r0), c1 = cv_mul(step, r1), c2 = cv_mul(step, r2),
                       c3 = cv_mul(step, r3);
                double v0 = cv_mul(v, c0);
                double v1 = cv_mul(v0, c1);
                double v2 = cv_mul(v1, c2);
                v = cv_mul(v2, c3);
                sp[0] = v0;
                sp[dt * CV_SDS] = v1;
                sp[2 * dt * CV_SDS] = v2;
                sp[3 * dt * CV_SDS] = v;
                tp += 4 * dt;
                sp += 4 * dt * CV_SDS;
            }
            for (; k < n1; k++) {
                v = cv_mul(v, cv_mul(step, *tp));
                *sp = v;
                tp += dt;
                sp += dt * CV_SDS;
            }
        }
        if (!down) { /* the rows below an interior starting row */
            const double *tp = M.row_dn + grow + rs - 1;
            double *sp = col + (rs - 1) * CV_SDS;
            double v = seed;
            for (int k = 0; k < rs; k++) {
                v = cv_mul(v, cv_mul(step_dn, *tp));
                *sp = v;
                tp -= 1;
                sp -= CV_SDS;
            }
        }
    }
}

/* One m8n8k4 FP64 MMA: D = A (8 x 4, row) * B (4 x 8, col) + D.  Lane l holds A[l >> 2][l & 3],
 * B[l & 3][l >> 2] and D[l >> 2][2 (l & 3) + {0, 1}]. */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void cv_dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
#endif

/* The accumulators of a lane: MMA tile (mt, nt), mt < 8 row tiles, nt < 2 column tiles, holds rows
 * 8 mt + r and columns 8 nt + 2 q + {0, 1} with r = lane >> 2, q = lane & 3; acc[4 mt + 2 nt + {0, 1}].
 * The 16 terms of the half-tile are contracted 4 at a time; lane q supplies the terms 2q and 2q+1 of
 * each group of 8 (one 16-byte read) to two consecutive MMAs.  NA = number of live row tiles.
 * On the host (test emulation) the same sums are formed with scalar FMAs in the same term order. */
template <int NA>
CV_HD void cv_w_fma(int lane, const CvWarpFixed &F, double *acc)
{
    const int r = lane >> 2, q = lane & 3;
#if defined(__CUDA_ARCH__)
    const double *sd = F.SD + r * CV_SDS + 2 * q;
    const double *pw = F.PW + r * CV_PWS + 2 * q;
#pragma unroll
    for (int k0 = 0; k0 < CV_HT; k0 += 8) {
        cv_pair b0 = cv_ld2(pw + k0), b1 = cv_ld2(pw + 8 * CV_PWS + k0);
#pragma unroll
        for (int mt = 0; mt < NA; mt++) {
            cv_pair a = cv_ld2(sd + mt * 8 * CV_SDS + k0);
            cv_dmma(acc[4 * mt + 0], acc[4 * mt + 1], a.x, b0.x);
            cv_dmma(acc[4 * mt + 2], acc[4 * mt + 3], a.x, b1.x);
            cv_dmma(acc[4 * mt + 0], acc[4 * mt + 1], a.y, b0.y);
            cv_dmma(acc[4 * mt + 2], acc[4 * mt + 3], a.y, b1.y);
        }
    }
#else
    for (int k0 = 0; k0 < CV_HT; k0 += 8)
        for (int mt = 0; mt < NA; mt++)
            for (int par = 0; par < 2; par++)     /* the .x terms, then the .y terms */
                for (int nt = 0; nt < 2; nt++)
                    for (int c = 0; c < 2; c++) {
                        const double *srow = F.SD + (8 * mt + r) * CV_SDS + k0 + par;
                        const double *pcol = F.PW + (8 * nt + 2 * q + c) * CV_PWS + k0 + par;
                        double v = acc[4 * mt + 2 * nt + c];
                        for (int kk = 0; kk < 4; kk++)
                            v = cv_fma(srow[2 * kk], pcol[2 * kk], v);
                        acc[4 * mt + 2 * nt + c] = v;
                    }
#endif
}

/* The accumulators of a lane go to the seed matrix (free after the last fma of a block) in slot
 * order, slot = row * 16 + column; the 16-byte chunks of odd rows are swapped between the halves of
 * the row so that a quarter warp (2 rows x 4 chunks) does not collide. */
CV_HD int cv_spill_index(int slot)
{
    return slot ^ ((slot & CV_W) >> 1); /* column ^= 8 on odd rows */
}
CV_HD void cv_w_spill(int lane, CvWarpFixed &F, const double *acc)
{
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            int slot = (8 * mt + r) * CV_W + 8 * nt + 2 * q;
            cv_st2(F.SD + cv_spill_index(slot), acc[4 * mt + 2 * nt], acc[4 * mt + 2 * nt + 1]);
        }
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

/* models.py:100-107 per bin: lane l finishes the slots l, l + 32, ... of the block (a compact
 * loop: the unrolled form of this, with its 32 logarithms, does not fit the instruction cache). */
CV_HD void cv_w_epilogue(int lane, const CvModelDesc &m, int blk, int nrows_blk, const CvWarpFixed &F,
                         CvPartial &part, double *out_probs)
{
    const CvTables &T = m.tab;
    const double *mult = T.slot_mult + blk * (CV_RB * CV_W);
    const double *cnt = T.slot_h + blk * (CV_RB * CV_W);
    const int *bin = T.slot_bin + blk * (CV_RB * CV_W);
    const int nslots = nrows_blk * CV_W;
#pragma unroll 2
    for (int e = lane; e < nslots; e += 32) {
        double mu = mult[e];
        if (mu == 0.0)
            continue; /* a bin that is not in hist */
        double p = cv_mul(F.SD[cv_spill_index(e)], mu);
        if (out_probs)
            out_probs[bin[e]] = p;
        cv_partial_add_mass(part, p);
        double h = cnt[e];
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

/* number of live 8-row groups of a block, rounded up to the instantiated 1, 2, 4, 8 */
CV_HD int cv_row_groups(int nrows_blk)
{
    int na = (nrows_blk + 7) >> 3;
    return na <= 1 ? 1 : na <= 2 ? 2 : na <= 4 ? 4 : 8;
}
