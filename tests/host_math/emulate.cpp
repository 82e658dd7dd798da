/* Test-only host execution of the kernel's per-point phases (covest_b200/csrc/cvpoint.h).
 *
 * The sm_100a kernel runs every phase with the 32 lanes of one warp and a __syncwarp() between
 * phases; here the same phase functions are called in a serial loop over `lane`, phase by phase,
 * which performs the same arithmetic in the same order (ballot, shuffles and the final shuffle tree
 * are plain loops over a 32-entry array).  This lets the CPU test-suite check the *formulation* --
 * scaling, seeds, recurrences, tables -- against the oracle without a GPU.  It is NOT a CPU
 * fallback: nothing under covest_b200/ loads it.
 *
 *   g++ -O2 -ffp-contract=off -fPIC -shared emulate.cpp -o libcv_emulate.so
 */
#include <cstring>
#include <vector>

#include "../../covest_b200/csrc/cvtables.h"

template <int NA>
static void emulate_blocks(const CvModelDesc &m, CvWarpMem &M, CvPartial *part, double *out_probs)
{
    CvWarpFixed &fx = *M.fx;
    const int S = m.n_err;
    const int cpg = cv_copies_per_group(S);
    std::vector<double> acc(32 * 4 * NA);
    for (int blk = 0; blk < m.n_blocks; blk++) {
        CvLaneGroup G[32];
        for (int lane = 0; lane < 32; lane++)
            G[lane] = cv_lane_group(lane, m, blk, M);
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int first = 1;; first += 32) {
            double b[32];
            int nlive = 32;
            bool any = false;
            for (int lane = 31; lane >= 0; lane--)
                if (cv_w_copy_pass(lane, m, M, first, &b[lane])) {
                    nlive = lane;
                    any = true;
                }
            for (int g = 0; g < nlive; g += cpg) {
                const int ncop = std::min(cpg, nlive - g);
                const int nterms = ncop * S;
                for (int lane = 0; lane < 32; lane++)
                    cv_w_mass(lane, m, first + g, nterms, S, M);
                for (int sub = 0; sub < nterms; sub += CV_CT) {
                    for (int lane = 0; lane < 32; lane++) {
                        int src = g + (sub + lane) / S;
                        CvTerm tm = cv_w_term(lane, m, first + g, nterms, S, sub, b[src < 31 ? src : 31], M);
                        cv_w_prep<NA>(lane, m, blk, tm, M);
                    }
                    const int nkg = (std::min(CV_CT, nterms - sub) + 3) >> 2;
                    for (int lane = 0; lane < 32; lane++)
                        cv_w_fused<NA>(lane, G[lane], 0, nkg, fx, acc.data() + 4 * NA * lane);
                }
            }
            if (any)
                break;
        }
        for (int lane = 0; lane < 32; lane++)
            cv_w_spill<NA>(lane, fx, acc.data() + 4 * NA * lane);
        for (int lane = 0; lane < 32; lane++)
            cv_w_epilogue<NA>(lane, m, blk, std::min(CV_GB, m.n_groups - blk * CV_GB), fx, part[lane],
                              out_probs);
    }
}

static void emulate_point(const CvModelDesc &m, const double *row_in, int clip, double *out_ll,
                          double *out_probs)
{
    static CvWarpFixed fx;
    const int S = m.n_err;
    std::vector<double> var(cv_warp_var_doubles(S));
    CvWarpMem M;
    cv_warp_mem_carve(M, &fx, var.data(), S);
    M.grp = m.tab.grp;
    double row[CV_MAX_PARAMS] = {0, 0, 0, 0, 0};
    for (int i = 0; i < m.n_param; i++)
        row[i] = row_in[i];
    for (int lane = 0; lane < 32; lane++)
        cv_w_header(lane, m, row, clip, M);
    CvPartial part[32];
    for (int lane = 0; lane < 32; lane++)
        part[lane] = CvPartial{0, 0, 0};
    switch (m.na) {
    case 1: emulate_blocks<1>(m, M, part, out_probs); break;
    case 2: emulate_blocks<2>(m, M, part, out_probs); break;
    case 4: emulate_blocks<4>(m, M, part, out_probs); break;
    default: emulate_blocks<8>(m, M, part, out_probs); break;
    }
    /* __shfl_down_sync tree: a lane whose source is out of range receives its own value */
    for (int d = 16; d >= 1; d >>= 1) {
        CvPartial prev[32];
        memcpy(prev, part, sizeof(prev));
        for (int lane = 0; lane < 32; lane++)
            cv_partial_merge(part[lane], prev[lane + d < 32 ? lane + d : lane]);
    }
    *out_ll = cv_point_finish(m, part[0]);
}

extern "C" int emu_loglik_batch(int model_kind, int k, int r, int n_err, int n_bins,
                                const int *bin_j, const double *bin_h, double tail,
                                const double *comb, const double *pow3, double threshold,
                                const double *bounds, long n_points, const double *par, int clip,
                                double *out_ll, double *out_probs)
{
    if (n_err > CV_MAX_ERR || n_err < 1)
        return -1;
    CvHostTables T;
    if (!cv_build_tables(n_bins, bin_j, bin_h, T).empty())
        return -2;
    CvModelDesc m;
    memset(&m, 0, sizeof(m));
    m.model_kind = model_kind;
    m.k = k;
    m.r = r;
    m.n_err = n_err;
    m.n_param = model_kind ? 5 : 2;
    m.n_bins = n_bins;
    m.n_groups = T.n_groups;
    m.na = T.na;
    m.n_blocks = T.n_blocks;
    m.max_bin = T.max_bin;
    m.tail = tail;
    m.threshold = threshold;
    for (int i = 0; i < CV_MAX_PARAMS; i++) {
        m.lo[i] = i < m.n_param ? bounds[2 * i] : NAN;
        m.hi[i] = i < m.n_param ? bounds[2 * i + 1] : NAN;
    }
    for (int s = 0; s < n_err; s++) {
        m.comb[s] = comb[s];
        m.pow3[s] = pow3[s];
    }
    m.tab = cv_tables_view(T);
    for (long p = 0; p < n_points; p++)
        emulate_point(m, par + (size_t)p * m.n_param, clip, out_ll + p,
                      out_probs ? out_probs + (size_t)p * n_bins : nullptr);
    return 0;
}
