/* Test-only host execution of the kernel's per-point phases (covest_b200/csrc/cvpoint.h).
 *
 * The sm_100a kernel runs every phase with 256 threads and a __syncthreads() between phases; here
 * the same phase functions are called in a serial loop over `tid`, phase by phase, which performs
 * the same arithmetic in the same order (the final cross-thread reduction is a plain loop).  This
 * lets the CPU test-suite check the *formulation* -- scaling, seeds, recurrences, tables -- against
 * the oracle without a GPU.  It is NOT a CPU fallback: nothing under covest_b200/ loads it.
 *
 *   g++ -O2 -ffp-contract=off -fPIC -shared emulate.cpp -o libcv_emulate.so
 */
#include <cstring>
#include <vector>

#include "../../covest_b200/csrc/cvtables.h"

static void emulate_point(const CvModelDesc &m, const double *row, int clip, double *out_ll,
                          double *out_probs)
{
    static CvPointShared sh;
    std::vector<double> acc((size_t)CV_NT * 32);
    for (int tid = 0; tid < CV_NT; tid++)
        cv_phase_header(tid, m, row, clip, sh);
    if (m.model_kind) {
        for (int first = 1; first < m.max_bin; first += CV_NT) {
            int best = sh.o_end;
            for (int tid = 0; tid < CV_NT; tid++)
                best = std::min(best, cv_phase_cut_candidate(tid, m, sh, first));
            sh.o_end = best;
            if (sh.o_end < first + CV_NT)
                break;
        }
    }
    int S = m.n_err, cpt = cv_copies_per_tile(S);
    CvPartial total = {0, 0, 0, 0};
    std::vector<CvPartial> part(CV_NT, CvPartial{0, 0, 0, 0});
    for (int blk = 0; blk < m.n_blocks; blk++) {
        int nrows_blk = std::min(CV_RB, m.n_rows - blk * CV_RB);
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int tile_o = 1; tile_o < sh.o_end; tile_o += cpt) {
            int ncop = std::min(cpt, sh.o_end - tile_o);
            int nterms = ncop * S;
            for (int tid = 0; tid < CV_NT; tid++)
                cv_phase_mass(tid, m, tile_o, nterms, sh);
            for (int tid = 0; tid < CV_NT; tid++)
                cv_phase_terms(tid, m, tile_o, nterms, sh);
            for (int tid = 0; tid < CV_NT; tid++)
                cv_phase_powers(tid, CV_NT, nterms, sh);
            for (int tid = 0; tid < CV_NT; tid++)
                cv_phase_seeds(tid, CV_NT, m, blk, nterms, sh);
            for (int tid = 0; tid < CV_NT; tid++)
                cv_phase_fma(tid, nterms, nrows_blk, sh, &acc[(size_t)tid * 32]);
        }
        for (int tid = 0; tid < CV_NT; tid++)
            cv_phase_spill(tid, sh, &acc[(size_t)tid * 32]);
        for (int tid = 0; tid < CV_NT; tid++)
            cv_phase_epilogue(tid, m, blk, sh, part[tid], out_probs);
    }
    for (int tid = 0; tid < CV_NT; tid++)
        cv_phase_publish(tid, sh, part[tid]);
    for (int lane = 0; lane < 32; lane++) {
        CvPartial q = cv_phase_fold(lane, sh);
        if (lane == 0)
            total = q;
        else
            cv_partial_merge(total, q);
    }
    *out_ll = cv_point_finish(m, total);
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
    m.n_rows = T.n_rows;
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
