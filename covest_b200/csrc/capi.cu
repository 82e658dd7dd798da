/*
 * capi.cu -- the C ABI of libcovest_b200.so (include/covest_b200.h): contexts, staging of host
 * buffers, kernel launches.  No arithmetic of the likelihood lives here -- and no CPU fallback:
 * every entry point needs a CUDA device.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/covest_b200.h"
#include "cvtables.h"
#include "factored.h"
#include "faithful.h"
#include "kernels.h"

#define CVB_CHUNK_POINTS (1LL << 22) /* staging granularity for host-resident batches */
#define CVB_MAX_TIMED_CHUNKS 64
#define CVB_TOPK_SORT_MIN 32768 /* batches from this size on select their top-K by a radix sort */

struct cvb_ctx {
    int device = 0;
    int n_sm = 0;
    int smem_max = 0; /* opt-in dynamic shared memory per CTA */
    CvModelDesc desc;
    std::vector<void *> owned; /* device allocations that live as long as the context */
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; /* device -> host copy of the values while the top-K runs */
    cudaEvent_t copy_ev = nullptr;
    cudaEvent_t copy_done = nullptr; /* the side copy has finished: the main stream waits for it */
    /* calls of one context share its scratch (work counters, plan, profiles, staging): a call on a
     * different stream than the previous one first waits for the previous call's work */
    cudaEvent_t order_ev = nullptr;
    cudaStream_t last_stream = nullptr;
    bool order_valid = false;
    /* growable staging */
    double *d_params = nullptr;
    size_t cap_params = 0; /* doubles */
    double *d_ll = nullptr;
    size_t cap_ll = 0;
    double *d_probs = nullptr;
    size_t cap_probs = 0;
    double *d_cand_ll = nullptr, *d_sel_ll = nullptr, *d_rows = nullptr;
    long long *d_cand_idx = nullptr, *d_sel_idx = nullptr;
    int cap_k = 0;
    int topk_ctas = 0;
    unsigned char *d_topk_scratch = nullptr;
    size_t cap_topk_scratch = 0;
    double *d_axes = nullptr;
    size_t cap_axes = 0;
    unsigned long long *d_counter = nullptr;
    double *d_sink = nullptr;
    /* timing */
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    int timed_chunks = 0;
    int last_launches = 0;
    cudaStream_t timed_stream = nullptr;
    /* the factored path of the repeats model (factored.h) */
    CvFactorWork fw;
    CvfSlots slots;            /* the row layout of the profiles (factored.h) */
    int table_lines = 0;       /* 64-slot lines of the histogram tables */
    const double *d_log_tab = nullptr;
    int path_mode = 0;         /* 0 auto, 1 per-point kernel only, 2 factored whenever supported */
    size_t w_limit = (size_t)2 << 30; /* doubles: 16 GiB of profiles per group range */
    double min_group = 12.0;   /* auto: points per (c, e) group below which the per-point kernel runs */
    long long min_points = 2048; /* auto: batches below this go straight to the per-point kernel */
    int last_path = 0;         /* 1 per-point, 2 factored (GEMM), 3 factored (prefix kernel) */
    double min_run = 4.0;      /* auto: points per q-run below which the GEMM runs instead of the prefix kernel */
    /* re-evaluation of marked points (faithful.h) */
    CvFaithTables faith = {nullptr, nullptr, 0, 0, nullptr, 0};
    unsigned long long *d_fixed = nullptr; /* two words: marked points of the most recent evaluation, work cursor */
    unsigned int *d_marked = nullptr;      /* their indices */
    size_t cap_marked = 0;
    std::string err;
};

static thread_local std::string g_create_error;

static int fail(cvb_ctx *ctx, int code, const std::string &msg)
{
    if (ctx)
        ctx->err = msg;
    else
        g_create_error = msg;
    return code;
}

static int fail_cuda(cvb_ctx *ctx, cudaError_t e, const char *what)
{
    return fail(ctx, CVB_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(call, what)                        \
    do {                                      \
        cudaError_t e_ = (call);              \
        if (e_ != cudaSuccess)                \
            return fail_cuda(ctx, e_, what);  \
    } while (0)

static bool is_device_ptr(const void *p)
{
    if (!p)
        return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

template <typename T>
static cudaError_t grow(T **buf, size_t *cap, size_t want)
{
    if (want <= *cap)
        return cudaSuccess;
    if (*buf)
        cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc((void **)buf, want * sizeof(T));
    if (e == cudaSuccess)
        *cap = want;
    return e;
}

template <typename T>
static cudaError_t upload(cvb_ctx *ctx, const std::vector<T> &v, const T **out)
{
    void *d = nullptr;
    size_t bytes = (v.empty() ? 1 : v.size()) * sizeof(T);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess)
        return e;
    ctx->owned.push_back(d);
    if (!v.empty())
        e = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = (const T *)d;
    return e;
}

extern "C" const char *cvb_version(void) { return "covest_b200 0.2 (sm_100a)"; }

extern "C" int cvb_abi_version(void) { return CVB_ABI_VERSION; }

extern "C" const char *cvb_last_error(const cvb_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int cvb_n_param(const cvb_ctx *ctx) { return ctx ? ctx->desc.n_param : CVB_EINVAL; }

extern "C" int cvb_device_sm_count(const cvb_ctx *ctx) { return ctx ? ctx->n_sm : CVB_EINVAL; }

extern "C" void cvb_ctx_destroy(cvb_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    for (void *p : ctx->owned)
        cudaFree(p);
    void *bufs[] = {ctx->d_params, ctx->d_ll,      ctx->d_probs, ctx->d_cand_ll, ctx->d_sel_ll,
                    ctx->d_rows,   ctx->d_cand_idx, ctx->d_sel_idx, ctx->d_axes,   ctx->d_counter,
                    ctx->d_sink,   ctx->d_topk_scratch, ctx->d_marked};
    for (void *p : bufs)
        if (p)
            cudaFree(p);
    for (cudaEvent_t e : ctx->ev)
        cudaEventDestroy(e);
    cvf_release(ctx->fw);
    if (ctx->copy_ev)
        cudaEventDestroy(ctx->copy_ev);
    if (ctx->copy_done)
        cudaEventDestroy(ctx->copy_done);
    if (ctx->order_ev)
        cudaEventDestroy(ctx->order_ev);
    if (ctx->copy_stream)
        cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int cvb_ctx_create(int model_kind, int k, int r, int max_error, int n_bins,
                              const int32_t *bin_j, const double *bin_h, double tail,
                              double threshold, const double *bounds, const double *comb,
                              const double *pow3, int device, cvb_ctx **out_ctx)
{
    cvb_ctx *ctx = nullptr; /* errors before the context exists go to the thread-local slot */
    if (!out_ctx)
        return fail(ctx, CVB_EINVAL, "out_ctx is NULL");
    *out_ctx = nullptr;
    if (model_kind != CVB_MODEL_BASIC && model_kind != CVB_MODEL_REPEATS)
        return fail(ctx, CVB_EINVAL, "model_kind must be 0 (basic) or 1 (repeats)");
    if (k < 1 || r < k)
        return fail(ctx, CVB_EINVAL, "need 1 <= k <= r");
    if (max_error < 1 || max_error > k + 1 || max_error > CV_MAX_ERR)
        return fail(ctx, CVB_EINVAL, "max_error must be in [1, min(k+1, 64)]");
    if (!bin_j || !bin_h || !bounds)
        return fail(ctx, CVB_EINVAL, "bin_j, bin_h and bounds are required");
    CvHostTables T;
    std::string why = cv_build_tables(n_bins, bin_j, bin_h, T);
    if (!why.empty())
        return fail(ctx, CVB_EINVAL, why);

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(ctx, CVB_ECUDA,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                        " (libcovest_b200 has no CPU path)");
    }
    if (device < 0 || device >= n_dev)
        return fail(ctx, CVB_EINVAL, "device ordinal out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return fail_cuda(ctx, e, "cudaSetDevice");

    cvb_ctx *c = new cvb_ctx();
    c->device = device;
    CvModelDesc &m = c->desc;
    memset(&m, 0, sizeof(m));
    m.model_kind = model_kind;
    m.k = k;
    m.r = r;
    m.n_err = max_error;
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
    for (int s = 0; s < max_error; s++) {
        if (comb) {
            m.comb[s] = comb[s];
        } else { /* C(k,s) * 3^s, exact in long double for every k <= 63 */
            long double v = 1.0L;
            for (int i = 1; i <= s; i++)
                v = v * (long double)(k - s + i) / (long double)i;
            m.comb[s] = (double)(v * powl(3.0L, (long double)s));
        }
        m.pow3[s] = pow3 ? pow3[s] : (s == 0 ? 1.0 : pow(3.0, (double)-s));
    }

    int rc = CVB_OK;
    do {
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
            break;
        c->n_sm = prop.multiProcessorCount;
        c->smem_max = (int)prop.sharedMemPerBlockOptin;
        if (prop.major < 10) {
            rc = fail(nullptr, CVB_ECUDA,
                      std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) +
                          std::to_string(prop.minor) + "); this library holds sm_100a code only");
            break;
        }
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess)
            break;
        CvTables &t = m.tab;
        if (cv_loglik_smem_bytes(m, c->smem_max) < 0) {
            rc = fail(nullptr, CVB_EINVAL, "histogram too long: its group tables do not fit the shared memory of an SM");
            break;
        }
        if ((e = upload(c, T.grp, &t.grp)) != cudaSuccess) break;
        if ((e = upload(c, T.slot_mult, &t.slot_mult)) != cudaSuccess) break;
        {
            std::vector<double> pr(T.slot_mult.size(), 0.0); /* the order of the profile kernel's stores */
            for (size_t line = 0; line + 64 <= T.slot_mult.size(); line += 64)
                for (int L = 0; L < 32; L++) {
                    const size_t s0 = line + 16 * (L >> 3) + (L & 7);
                    pr[line + 2 * L] = T.slot_mult[s0];
                    pr[line + 2 * L + 1] = T.slot_mult[s0 + 8];
                }
            if ((e = upload(c, pr, &t.slot_mult_pair)) != cudaSuccess) break;
        }
        if ((e = upload(c, T.slot_h, &t.slot_h)) != cudaSuccess) break;
        if ((e = upload(c, T.slot_bin, &t.slot_bin)) != cudaSuccess) break;
        if ((e = upload(c, T.copy_log_h, &t.copy_log_h)) != cudaSuccess) break;
        if ((e = upload(c, T.copy_log_l, &t.copy_log_l)) != cudaSuccess) break;
        if ((e = upload(c, T.run_first, &t.run_first)) != cudaSuccess) break;
        if ((e = upload(c, T.run_len, &t.run_len)) != cudaSuccess) break;
        if ((e = upload(c, T.blk_run_begin, &t.blk_run_begin)) != cudaSuccess) break;
        {
            /* rows of the profiles: the lines with counts, the others summed (COVEST_B200_ROWS=full: every line) */
            const char *rows = getenv("COVEST_B200_ROWS");
            std::vector<int> line_map;
            std::vector<double2> mh;
            std::vector<double> row_h;
            cvf_build_slots(T.slot_mult, T.slot_h, !(rows && !strcmp(rows, "full")), line_map, mh, row_h,
                            &c->slots.sum_line);
            c->table_lines = (int)line_map.size();
            c->slots.nsteps = (int)(row_h.size() / 64);
            if ((e = upload(c, line_map, &c->slots.line_map)) != cudaSuccess) break;
            if ((e = upload(c, mh, &c->slots.slot_mh)) != cudaSuccess) break;
            if ((e = upload(c, cvf_step_masks(row_h), &c->slots.step_mask)) != cudaSuccess) break;
            c->slots.counts_first = cvf_counts_first(row_h);
            std::vector<double> lt(2 * CV_LOG_N);
            cv_log_table(lt.data());
            if ((e = upload(c, lt, &c->d_log_tab)) != cudaSuccess) break;
        }
        if (const char *pm = getenv("COVEST_B200_PATH"))
            c->path_mode = !strcmp(pm, "direct") ? 1 : !strcmp(pm, "factored") ? 2 : !strcmp(pm, "gemm") ? 3
                           : !strcmp(pm, "prefix") ? 4 : !strcmp(pm, "faithful") ? 5 : 0;
        if (const char *pv = getenv("COVEST_B200_PREFIX_KERNEL")) /* which prefix kernel (factored.cu, cvf_eval) */
            c->fw.prefix_version = atoi(pv);
        if (const char *to = getenv("COVEST_B200_TILE_ORDER")) /* 0 = by descending cost, 1 = group by group */
            c->fw.tile_interleave = atoi(to) != 0;
        if (const char *wl = getenv("COVEST_B200_PROFILE_MIB"))
            if (atoll(wl) > 0)
                c->w_limit = (size_t)atoll(wl) * (1 << 20) / sizeof(double);
        { /* tables of the term-by-term re-evaluation (faithful.h): by bin index j - 1 */
            const int j_all = T.max_bin > 0 ? T.max_bin : 1;
            std::vector<double> cnt(j_all, -1.0), rcp(j_all);
            int j_counted = 0;
            for (int b = 0; b < n_bins; b++) {
                const int j = (int)bin_j[b];
                if (j >= 1 && j <= j_all) {
                    cnt[j - 1] = bin_h[b];
                    if (bin_h[b] != 0.0 && j > j_counted)
                        j_counted = j;
                }
            }
            for (int j = 1; j <= j_all; j++)
                rcp[j - 1] = 1.0 / (double)j;
            if ((e = upload(c, cnt, &c->faith.cnt_of_j)) != cudaSuccess) break;
            if ((e = upload(c, rcp, &c->faith.rcp)) != cudaSuccess) break;
            c->faith.j_all = j_all;
            c->faith.j_counted = j_counted;
            c->faith.acc_doubles = ((long long)j_all + 31) & ~31LL;
            void *scr = nullptr;
            if ((e = cudaMalloc(&scr, (size_t)cv_faithful_warps(c->n_sm) * (size_t)c->faith.acc_doubles * sizeof(double))) != cudaSuccess) break;
            c->owned.push_back(scr);
            c->faith.scratch = (double *)scr;
            if ((e = cudaMalloc((void **)&c->d_fixed, 4 * sizeof(unsigned long long))) != cudaSuccess) break;
            c->owned.push_back(c->d_fixed);
        }
        if ((e = cudaMalloc((void **)&c->d_counter, sizeof(unsigned long long))) != cudaSuccess) break;
        if ((e = cudaMalloc((void **)&c->d_sink, 64)) != cudaSuccess) break;
    } while (0);
    if (rc == CVB_OK && e != cudaSuccess)
        rc = fail_cuda(nullptr, e, "context setup");
    if (rc != CVB_OK) {
        cvb_ctx_destroy(c);
        return rc;
    }
    *out_ctx = c;
    return CVB_OK;
}

/* ---- timing ------------------------------------------------------------------------------- */
extern "C" int cvb_set_timing(cvb_ctx *ctx, int enabled)
{
    if (!ctx)
        return CVB_EINVAL;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (enabled && ctx->ev.empty()) {
        ctx->ev.resize(2 * CVB_MAX_TIMED_CHUNKS);
        for (auto &e : ctx->ev)
            CU(cudaEventCreate(&e), "cudaEventCreate");
    }
    ctx->timing = enabled != 0;
    ctx->timed_chunks = 0;
    return CVB_OK;
}

extern "C" int cvb_last_kernel_ms(cvb_ctx *ctx, double *out_ms, int *out_launches)
{
    if (!ctx)
        return CVB_EINVAL;
    if (out_launches)
        *out_launches = ctx->last_launches;
    if (out_ms) {
        *out_ms = 0.0;
        if (!ctx->timing)
            return fail(ctx, CVB_EINVAL, "timing is off (cvb_set_timing)");
        CU(cudaSetDevice(ctx->device), "cudaSetDevice");
        for (int i = 0; i < ctx->timed_chunks; i++) {
            float ms = 0.f;
            CU(cudaEventSynchronize(ctx->ev[2 * i + 1]), "cudaEventSynchronize");
            CU(cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]), "cudaEventElapsedTime");
            *out_ms += ms;
        }
    }
    return CVB_OK;
}

/* one evaluation of n points on the device, optionally bracketed by events: the factored path
 * when the batch is large and groups well (never when per-bin probabilities are wanted), else the
 * per-point kernel */
static int launch_loglik(cvb_ctx *ctx, const CvLattice &lat, const double *const *lat_axes_host,
                         const double *d_params, long long n, int clip, double *d_ll, double *d_probs,
                         cudaStream_t s)
{
    bool timed = ctx->timing && ctx->timed_chunks < CVB_MAX_TIMED_CHUNKS;
    if (timed)
        CU(cudaEventRecord(ctx->ev[2 * ctx->timed_chunks], s), "cudaEventRecord");
    int used = 0;
    const bool forced = ctx->path_mode >= 2;
    const bool faithful_only = ctx->path_mode == 5 && !d_probs;
    const bool try_factored = !d_probs && ctx->path_mode != 1 && !faithful_only && cvf_supported(ctx->desc) &&
                              (forced || n >= ctx->min_points);
    if (faithful_only) {
        CU(cv_launch_mark_all(d_ll, n, s), "cv_mark_all_kernel launch");
        ctx->last_launches++;
        used = 3;
    }
    if (try_factored) {
        ctx->fw.timed = ctx->timing;
        CU(cvf_eval(ctx->desc, lat, lat_axes_host, d_params, n, clip, d_ll, ctx->slots,
                    ctx->d_log_tab, ctx->fw, ctx->n_sm,
                    ctx->smem_max, ctx->w_limit, forced ? 0.0 : ctx->min_group, ctx->min_run,
                    ctx->path_mode == 3 ? 1 : ctx->path_mode == 4 ? 2 : 0, s, &used),
           "factored evaluation");
        ctx->last_launches += ctx->fw.launches;
    }
    if (!used) {
        CU(cv_launch_loglik(ctx->desc, lat, d_params, n, clip, d_ll, d_probs, ctx->d_counter, ctx->n_sm,
                            ctx->smem_max, s),
           "cv_loglik_kernel launch");
        ctx->last_launches++;
    }
    ctx->last_path = used ? 1 + used : 1;
    /* points whose value hinges on the reference's subnormal roundings: again, term by term */
    CU(grow(&ctx->d_marked, &ctx->cap_marked, 2 * (size_t)n), "cudaMalloc(marked points)");
    CU(cv_launch_faithful(ctx->desc, lat, d_params, n, clip, d_ll, ctx->faith, ctx->n_sm, ctx->d_marked,
                          ctx->d_fixed, s),
       "cv_faithful_kernel launch");
    ctx->last_launches += 3; /* the scan for marked points, their re-evaluation (warp / CTA per point) */
    if (timed) {
        CU(cudaEventRecord(ctx->ev[2 * ctx->timed_chunks + 1], s), "cudaEventRecord");
        ctx->timed_chunks++;
    }
    return CVB_OK;
}

static int topk_device(cvb_ctx *ctx, const CvLattice &lat, const double *d_ll, const double *d_params,
                       long long n, int K, double *out_rows, cudaStream_t s);

/* stream ordering between the calls of one context (see cvb_ctx::order_ev) */
static int order_enter(cvb_ctx *ctx, cudaStream_t s)
{
    if (ctx->order_valid && ctx->last_stream != s)
        CU(cudaStreamWaitEvent(s, ctx->order_ev, 0), "cudaStreamWaitEvent");
    return CVB_OK;
}

static int order_leave(cvb_ctx *ctx, cudaStream_t s)
{
    if (!ctx->order_ev)
        CU(cudaEventCreateWithFlags(&ctx->order_ev, cudaEventDisableTiming), "cudaEventCreate");
    CU(cudaEventRecord(ctx->order_ev, s), "cudaEventRecord");
    ctx->last_stream = s;
    ctx->order_valid = true;
    return CVB_OK;
}

static int eval_batch(cvb_ctx *ctx, int64_t n_points, const double *params, int clip, double *out_p,
                      double *out_ll, int k_best, double *out_rows, void *stream)
{
    if (!ctx)
        return CVB_EINVAL;
    if (n_points < 0 || (n_points > 0 && !params))
        return fail(ctx, CVB_EINVAL, "n_points < 0 or params is NULL");
    if (!out_ll && !out_p && k_best <= 0)
        return fail(ctx, CVB_EINVAL, "no output buffer");
    if (k_best < 0 || (k_best > 0 && !out_rows))
        return fail(ctx, CVB_EINVAL, "k_best > 0 needs out_rows");
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    ctx->timed_chunks = 0;
    ctx->last_launches = 0;
    if (n_points == 0 && k_best <= 0)
        return CVB_OK;
    if (int rc = order_enter(ctx, s))
        return rc;
    const int np = ctx->desc.n_param;
    const long long nb = ctx->desc.n_bins;
    const bool p_dev = is_device_ptr(params);
    const bool l_dev = out_ll ? is_device_ptr(out_ll) : true;
    const bool q_dev = out_p ? is_device_ptr(out_p) : true;
    const bool all_dev = p_dev && l_dev && q_dev;
    CvLattice lat;
    memset(&lat, 0, sizeof(lat));

    /* the row gather of a top-K needs the whole batch resident: no chunking then */
    long long chunk = (all_dev || k_best > 0) ? n_points : CVB_CHUNK_POINTS;
    if (chunk < 1)
        chunk = 1;
    if (out_p && !q_dev) { /* keep the probability staging buffer below ~1 GiB */
        long long lim = (1LL << 27) / (nb > 0 ? nb : 1);
        if (lim < 1)
            lim = 1;
        if (chunk > lim)
            chunk = lim;
    }
    if (chunk > n_points)
        chunk = n_points;
    if (!p_dev)
        CU(grow(&ctx->d_params, &ctx->cap_params, (size_t)chunk * np), "cudaMalloc(params staging)");
    if (!l_dev || !out_ll)
        CU(grow(&ctx->d_ll, &ctx->cap_ll, (size_t)chunk), "cudaMalloc(loglik staging)");
    const double *dp_all = p_dev ? params : ctx->d_params;
    const double *dl_all = (out_ll && l_dev) ? out_ll : ctx->d_ll;
    if (out_p && !q_dev)
        CU(grow(&ctx->d_probs, &ctx->cap_probs, (size_t)chunk * nb), "cudaMalloc(probs staging)");

    bool side_copy = false;
    for (long long off = 0; off < n_points; off += chunk) {
        long long n = n_points - off < chunk ? n_points - off : chunk;
        const double *dp = params + off * np;
        if (!p_dev) {
            CU(cudaMemcpyAsync(ctx->d_params, dp, (size_t)n * np * sizeof(double),
                               cudaMemcpyHostToDevice, s),
               "cudaMemcpyAsync(params)");
            dp = ctx->d_params;
        }
        double *dl = (out_ll && l_dev) ? out_ll + off : ctx->d_ll;
        double *dq = out_p ? (q_dev ? out_p + off * nb : ctx->d_probs) : nullptr;
        int rc = launch_loglik(ctx, lat, nullptr, dp, n, clip, dl, dq, s);
        if (rc != CVB_OK)
            return rc;
        if (out_ll && !l_dev) {
            /* with a top-K to follow the values leave on a stream of their own, next to it */
            cudaStream_t cs = s;
            if (k_best > 0 && n >= (1 << 16)) {
                if (!ctx->copy_stream)
                    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
                if (!ctx->copy_ev)
                    CU(cudaEventCreateWithFlags(&ctx->copy_ev, cudaEventDisableTiming), "cudaEventCreate");
                if (!ctx->copy_done)
                    CU(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming), "cudaEventCreate");
                CU(cudaEventRecord(ctx->copy_ev, s), "cudaEventRecord");
                CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev, 0), "cudaStreamWaitEvent");
                cs = ctx->copy_stream;
                side_copy = true;
            }
            CU(cudaMemcpyAsync(out_ll + off, dl, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, cs),
               "cudaMemcpyAsync(loglik)");
            if (cs != s)
                CU(cudaEventRecord(ctx->copy_done, cs), "cudaEventRecord");
        }
        if (out_p && !q_dev)
            CU(cudaMemcpyAsync(out_p + off * nb, dq, (size_t)n * nb * sizeof(double),
                               cudaMemcpyDeviceToHost, s),
               "cudaMemcpyAsync(probs)");
        if (!all_dev && off + chunk < n_points)
            CU(cudaStreamSynchronize(s), "cudaStreamSynchronize"); /* staging is reused */
    }
    bool rows_to_host = false;
    if (k_best > 0) {
        int rc = topk_device(ctx, lat, dl_all, dp_all, n_points, k_best, out_rows, s);
        if (rc != CVB_OK)
            return rc;
        rows_to_host = !is_device_ptr(out_rows);
    }
    if (side_copy) /* the values left on the side stream: the main stream ends after them */
        CU(cudaStreamWaitEvent(s, ctx->copy_done, 0), "cudaStreamWaitEvent");
    if (int rc = order_leave(ctx, s))
        return rc;
    if (!all_dev || rows_to_host) /* one synchronisation per call, and only when something went to host memory */
        CU(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    return CVB_OK;
}

extern "C" int cvb_loglik_batch(cvb_ctx *ctx, int64_t n_points, const double *params, double *out_ll,
                                void *stream)
{
    if (ctx && !out_ll)
        return fail(ctx, CVB_EINVAL, "out_ll is NULL");
    return eval_batch(ctx, n_points, params, 1, nullptr, out_ll, 0, nullptr, stream);
}

extern "C" int cvb_probs_batch(cvb_ctx *ctx, int64_t n_points, const double *params, int clip,
                               double *out_p, double *out_ll, void *stream)
{
    if (ctx && !out_p)
        return fail(ctx, CVB_EINVAL, "out_p is NULL");
    return eval_batch(ctx, n_points, params, clip ? 1 : 0, out_p, out_ll, 0, nullptr, stream);
}

extern "C" int cvb_loglik_topk(cvb_ctx *ctx, int64_t n_points, const double *params, double *out_ll,
                               int k_best, double *out_rows, void *stream)
{
    if (ctx && (k_best < 1 || !out_rows))
        return fail(ctx, CVB_EINVAL, "cvb_loglik_topk needs k_best >= 1 and out_rows");
    return eval_batch(ctx, n_points, params, 1, nullptr, out_ll, k_best, out_rows, stream);
}

/* ---- top-K -------------------------------------------------------------------------------- */
static int ensure_topk(cvb_ctx *ctx, int K)
{
    if (K <= ctx->cap_k)
        return CVB_OK;
    void *old[] = {ctx->d_cand_ll, ctx->d_cand_idx, ctx->d_sel_ll, ctx->d_sel_idx, ctx->d_rows};
    for (void *p : old)
        if (p)
            cudaFree(p);
    ctx->d_cand_ll = ctx->d_sel_ll = ctx->d_rows = nullptr;
    ctx->d_cand_idx = ctx->d_sel_idx = nullptr;
    ctx->cap_k = 0;
    ctx->topk_ctas = 2 * ctx->n_sm;
    size_t nc = (size_t)ctx->topk_ctas * K;
    CU(cudaMalloc((void **)&ctx->d_cand_ll, nc * sizeof(double)), "cudaMalloc(topk)");
    CU(cudaMalloc((void **)&ctx->d_cand_idx, nc * sizeof(long long)), "cudaMalloc(topk)");
    CU(cudaMalloc((void **)&ctx->d_sel_ll, (size_t)K * sizeof(double)), "cudaMalloc(topk)");
    CU(cudaMalloc((void **)&ctx->d_sel_idx, (size_t)K * sizeof(long long)), "cudaMalloc(topk)");
    CU(cudaMalloc((void **)&ctx->d_rows, (size_t)K * (1 + CV_MAX_PARAMS) * sizeof(double)),
       "cudaMalloc(topk)");
    ctx->cap_k = K;
    return CVB_OK;
}

/* d_ll: device, n entries.  params: device or null (lattice). */
static int topk_device(cvb_ctx *ctx, const CvLattice &lat, const double *d_ll, const double *d_params,
                       long long n, int K, double *out_rows, cudaStream_t s)
{
    int rc = ensure_topk(ctx, K);
    if (rc != CVB_OK)
        return rc;
    if (n >= CVB_TOPK_SORT_MIN && cv_topk_select_fits(n, K) && !getenv("COVEST_B200_TOPK_SORT")) {
        /* large batch, few rows wanted: radix selection (topk.cu) */
        CU(grow(&ctx->d_topk_scratch, &ctx->cap_topk_scratch, cv_topk_select_bytes()), "cudaMalloc(topk scratch)");
        CU(cv_launch_topk_radix_select(d_ll, n, K, ctx->d_topk_scratch, ctx->n_sm, ctx->d_sel_ll, ctx->d_sel_idx, s),
           "top-K selection");
        ctx->last_launches += 1;
    } else if (n >= CVB_TOPK_SORT_MIN && n <= 0x7fffffffLL) { /* large batch: one radix sort (topk.cu) */
        const size_t need = cv_topk_sort_bytes(n);
        CU(grow(&ctx->d_topk_scratch, &ctx->cap_topk_scratch, need), "cudaMalloc(topk scratch)");
        CU(cv_launch_topk_sort(d_ll, n, K, ctx->d_topk_scratch, ctx->cap_topk_scratch, ctx->d_sel_ll,
                               ctx->d_sel_idx, s),
           "top-K sort");
        ctx->last_launches += 12; /* keys, the passes of the radix sort, take */
    } else {
        int ctas = ctx->topk_ctas;
        if ((long long)ctas * 1024 > n) /* small inputs: fewer, fuller slices */
            ctas = (int)((n + 1023) / 1024);
        if (ctas < 1)
            ctas = 1;
        CU(cv_launch_topk(d_ll, n, K, ctx->d_cand_ll, ctx->d_cand_idx, ctas, ctx->d_sel_ll,
                          ctx->d_sel_idx, s),
           "cv_topk_select launch");
        ctx->last_launches += 2;
    }
    const int np = ctx->desc.n_param;
    const bool r_dev = is_device_ptr(out_rows);
    double *dr = r_dev ? out_rows : ctx->d_rows;
    CU(cv_launch_gather_rows(lat, d_params, np, ctx->d_sel_ll, ctx->d_sel_idx, K, dr, s),
       "cv_gather_rows launch");
    ctx->last_launches += 1; /* the row gather */
    if (!r_dev) /* the caller synchronises the stream before it returns */
        CU(cudaMemcpyAsync(out_rows, dr, (size_t)K * (1 + np) * sizeof(double), cudaMemcpyDeviceToHost, s),
           "cudaMemcpyAsync(rows)");
    return CVB_OK;
}

extern "C" int cvb_topk(cvb_ctx *ctx, int64_t n_points, const double *ll, const double *params,
                        int k_best, double *out_rows, void *stream)
{
    if (!ctx)
        return CVB_EINVAL;
    if (n_points < 0 || k_best < 1 || !out_rows || (n_points > 0 && (!ll || !params)))
        return fail(ctx, CVB_EINVAL, "bad arguments to cvb_topk");
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    const int np = ctx->desc.n_param;
    ctx->last_launches = 0;
    if (int rc = order_enter(ctx, s))
        return rc;
    const double *dl = ll, *dp = params;
    if (n_points > 0 && !is_device_ptr(ll)) {
        CU(grow(&ctx->d_ll, &ctx->cap_ll, (size_t)n_points), "cudaMalloc(loglik staging)");
        CU(cudaMemcpyAsync(ctx->d_ll, ll, (size_t)n_points * sizeof(double), cudaMemcpyHostToDevice, s),
           "cudaMemcpyAsync(ll)");
        dl = ctx->d_ll;
    }
    if (n_points > 0 && !is_device_ptr(params)) {
        CU(grow(&ctx->d_params, &ctx->cap_params, (size_t)n_points * np), "cudaMalloc(params staging)");
        CU(cudaMemcpyAsync(ctx->d_params, params, (size_t)n_points * np * sizeof(double),
                           cudaMemcpyHostToDevice, s),
           "cudaMemcpyAsync(params)");
        dp = ctx->d_params;
    }
    CvLattice lat;
    memset(&lat, 0, sizeof(lat));
    if (int rc = topk_device(ctx, lat, dl, dp, n_points, k_best, out_rows, s))
        return rc;
    if (int rc = order_leave(ctx, s))
        return rc;
    if (!is_device_ptr(out_rows))
        CU(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    return CVB_OK;
}

/* ---- lattice ------------------------------------------------------------------------------ */
/* slices with host output from this size on are evaluated in parts (measured on cfg3, 10^6 points: two parts cost
 * 0.11 ms of a 1.7 ms call on one GPU -- shorter K1 / K2p launches leave SMs idle at their ends) */
#define CVB_LATTICE_PART_MIN ((int64_t)1 << 22)
extern "C" int cvb_lattice_eval(cvb_ctx *ctx, const int32_t *axis_len, const double *axis_values,
                                int64_t first, int64_t stride, int64_t block, int64_t count,
                                double *out_ll, int k_best, double *out_rows, void *stream)
{
    if (!ctx)
        return CVB_EINVAL;
    if (!axis_len || !axis_values || first < 0 || stride < 1 || block < 1 || count < 0)
        return fail(ctx, CVB_EINVAL, "bad lattice arguments");
    if (k_best < 0 || (k_best > 0 && !out_rows))
        return fail(ctx, CVB_EINVAL, "k_best > 0 needs out_rows");
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    const int np = ctx->desc.n_param;
    ctx->timed_chunks = 0;
    ctx->last_launches = 0;
    if (int rc = order_enter(ctx, s))
        return rc;
    size_t total_vals = 0;
    double total_pts = 1.0;
    for (int a = 0; a < np; a++) {
        if (axis_len[a] < 1)
            return fail(ctx, CVB_EINVAL, "empty lattice axis");
        total_vals += (size_t)axis_len[a];
        total_pts *= (double)axis_len[a];
    }
    if (count > 0) {
        const int64_t last_run = (count - 1) / block, last_in = (count - 1) % block;
        if (((double)first + (double)last_run * (double)stride) * (double)block + (double)last_in >= total_pts)
            return fail(ctx, CVB_EINVAL, "lattice slice runs past the end of the lattice");
    }
    CU(grow(&ctx->d_axes, &ctx->cap_axes, total_vals), "cudaMalloc(axes)");
    CU(cudaMemcpyAsync(ctx->d_axes, axis_values, total_vals * sizeof(double), cudaMemcpyHostToDevice, s),
       "cudaMemcpyAsync(axes)");
    CvLattice lat;
    memset(&lat, 0, sizeof(lat));
    lat.enabled = 1;
    lat.n_axes = np;
    size_t at = 0;
    for (int a = 0; a < np; a++) {
        lat.len[a] = axis_len[a];
        lat.axis[a] = ctx->d_axes + at;
        at += (size_t)axis_len[a];
    }
    lat.first = first;
    lat.stride = stride;
    lat.block = block;
    const double *axes_host[CV_MAX_PARAMS] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    at = 0;
    for (int a = 0; a < np; a++) {
        axes_host[a] = axis_values + at;
        at += (size_t)axis_len[a];
    }
    const bool l_dev = out_ll && is_device_ptr(out_ll);
    double *dl = l_dev ? out_ll : nullptr;
    if (!dl) {
        CU(grow(&ctx->d_ll, &ctx->cap_ll, (size_t)(count > 0 ? count : 1)), "cudaMalloc(loglik staging)");
        dl = ctx->d_ll;
    }
    bool need_sync = false, side_copy = false;
    /* A long slice whose values go to host memory is evaluated in parts of whole runs (and whole
     * (coverage, error_rate) groups), so that the values of a part travel while the next part is
     * evaluated: at 8 bytes per point the copy is otherwise up to a third of the call (cfg5: 100 MB per
     * rank; and the ranks of a box finish together and share the way to host memory). */
    int64_t part = 0;
    if (out_ll && !l_dev && count >= CVB_LATTICE_PART_MIN) {
        int64_t unit = block;
        if (block == 1 && stride == 1 && np == 5)
            unit = (int64_t)axis_len[2] * axis_len[3] * axis_len[4];
        /* parts of about 2^21 points, at most eight */
        int64_t n_parts = count >> 21;
        n_parts = n_parts < 2 ? 2 : n_parts > 8 ? 8 : n_parts;
        const int64_t want = (count + n_parts - 1) / n_parts;
        if (unit >= 1 && unit <= want && (block > 1 || first % unit == 0))
            part = (want + unit - 1) / unit * unit;
    }
    if (part > 0) {
        if (!ctx->copy_stream)
            CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!ctx->copy_ev)
            CU(cudaEventCreateWithFlags(&ctx->copy_ev, cudaEventDisableTiming), "cudaEventCreate");
        if (!ctx->copy_done)
            CU(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming), "cudaEventCreate");
        for (int64_t off = 0; off < count; off += part) {
            const int64_t n_part = count - off < part ? count - off : part;
            CvLattice lp = lat;
            lp.first = first + (off / block) * stride;
            int rc = launch_loglik(ctx, lp, axes_host, nullptr, n_part, 1, dl + off, nullptr, s);
            if (rc != CVB_OK)
                return rc;
            CU(cudaEventRecord(ctx->copy_ev, s), "cudaEventRecord");
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev, 0), "cudaStreamWaitEvent");
            CU(cudaMemcpyAsync(out_ll + off, dl + off, (size_t)n_part * sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->copy_stream),
               "cudaMemcpyAsync(loglik)");
        }
        CU(cudaEventRecord(ctx->copy_done, ctx->copy_stream), "cudaEventRecord");
        side_copy = true;
        need_sync = true;
    } else if (count > 0) {
        int rc = launch_loglik(ctx, lat, axes_host, nullptr, count, 1, dl, nullptr, s);
        if (rc != CVB_OK)
            return rc;
    }
    if (out_ll && !l_dev && count > 0 && part == 0) {
        cudaStream_t cs = s;
        if (k_best > 0 && count >= (1 << 16)) { /* the values leave next to the top-K, on a stream of their own */
            if (!ctx->copy_stream)
                CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
            if (!ctx->copy_ev)
                CU(cudaEventCreateWithFlags(&ctx->copy_ev, cudaEventDisableTiming), "cudaEventCreate");
            if (!ctx->copy_done)
                CU(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming), "cudaEventCreate");
            CU(cudaEventRecord(ctx->copy_ev, s), "cudaEventRecord");
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev, 0), "cudaStreamWaitEvent");
            cs = ctx->copy_stream;
            side_copy = true;
        }
        CU(cudaMemcpyAsync(out_ll, dl, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, cs),
           "cudaMemcpyAsync(loglik)");
        if (side_copy)
            CU(cudaEventRecord(ctx->copy_done, cs), "cudaEventRecord");
        need_sync = true;
    }
    if (k_best > 0) {
        int rc = topk_device(ctx, lat, dl, nullptr, count, k_best, out_rows, s);
        if (rc != CVB_OK)
            return rc;
        need_sync = need_sync || !is_device_ptr(out_rows);
    }
    if (side_copy)
        CU(cudaStreamWaitEvent(s, ctx->copy_done, 0), "cudaStreamWaitEvent");
    if (int rc = order_leave(ctx, s))
        return rc;
    if (need_sync)
        CU(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    return CVB_OK;
}

extern "C" int cvb_merge_rows(const double *rows, int n_rows, int n_cols, int k_best, double *out_rows,
                              void *stream)
{
    if (!rows || !out_rows || n_rows < 0 || n_rows > 2048 || n_cols < 1 || k_best < 1)
        return CVB_EINVAL;
    if (!is_device_ptr(rows) || !is_device_ptr(out_rows))
        return CVB_EINVAL;
    cudaError_t e = cv_launch_merge_rows(rows, n_rows, n_cols, k_best, out_rows, (cudaStream_t)stream);
    return e == cudaSuccess ? CVB_OK : CVB_ECUDA;
}

/* ---- path selection and facts about the last evaluation -------------------------------------- */
extern "C" int cvb_set_path(cvb_ctx *ctx, int mode)
{
    if (!ctx)
        return CVB_EINVAL;
    if (mode < 0 || mode > 5)
        return fail(ctx, CVB_EINVAL, "path mode must be 0 (auto), 1 (per-point), 2 (factored), 3 (factored, GEMM), "
                                     "4 (factored, prefix kernel) or 5 (term by term)");
    ctx->path_mode = mode;
    return CVB_OK;
}

extern "C" int cvb_last_path_info(cvb_ctx *ctx, double *out, int n_out)
{
    if (!ctx || !out || n_out < 1)
        return CVB_EINVAL;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    double v[CVB_PATH_INFO_LEN] = {0};
    v[0] = ctx->last_path;
    {
        unsigned long long fixed = 0; /* waits for the work queued on the device */
        CU(cudaMemcpy(&fixed, ctx->d_fixed, sizeof(fixed), cudaMemcpyDeviceToHost), "cudaMemcpy");
        v[9] = (double)fixed;
    }
    if (ctx->last_path >= 2) {
        v[10] = (double)ctx->fw.analytic;
        v[11] = (double)(ctx->slots.nsteps * 64 - (ctx->slots.sum_line >= 0 ? 32 : 0));
        const CvFactorWork &w = ctx->fw;
        v[1] = (double)w.n_groups;
        v[2] = (double)w.n_tiles;
        v[3] = (double)w.n_items;
        v[4] = (double)w.w_doubles;
        v[8] = (double)w.n_runs;
        if (w.timed && w.ev[0]) {
            float ms = 0.f;
            CU(cudaEventSynchronize(w.ev[3]), "cudaEventSynchronize");
            CU(cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]), "cudaEventElapsedTime");
            v[5] = ms; /* keys, sort, group tables (includes one host synchronisation) */
            CU(cudaEventElapsedTime(&ms, w.ev[1], w.ev[2]), "cudaEventElapsedTime");
            v[6] = ms; /* profiles (and the tiles of all but the last group range) */
            CU(cudaEventElapsedTime(&ms, w.ev[2], w.ev[3]), "cudaEventElapsedTime");
            v[7] = ms; /* tiles of the last group range */
        }
    }
    for (int i = 0; i < n_out && i < CVB_PATH_INFO_LEN; i++)
        out[i] = v[i];
    return CVB_OK;
}

/* ---- roofline probe ----------------------------------------------------------------------- */
extern "C" int cvb_fp64_peak(cvb_ctx *ctx, int kind, int reps, double *out_tflops)
{
    if (!ctx || !out_tflops || (kind < 0 || kind > 2))
        return ctx ? fail(ctx, CVB_EINVAL, "bad arguments to cvb_fp64_peak") : CVB_EINVAL;
    CU(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a), "cudaEventCreate");
    CU(cudaEventCreate(&b), "cudaEventCreate");
    double best = 0.0, flop = 0.0;
    if (reps < 1)
        reps = 1;
    for (int rep = 0; rep < reps + 1; rep++) { /* first pass warms up */
        CU(cudaEventRecord(a, ctx->stream), "cudaEventRecord");
        CU(cv_launch_peak_probe(kind, ctx->n_sm * 8, 4096, ctx->d_sink, &flop, ctx->stream),
           "cv_peak_probe launch");
        CU(cudaEventRecord(b, ctx->stream), "cudaEventRecord");
        CU(cudaEventSynchronize(b), "cudaEventSynchronize");
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b), "cudaEventElapsedTime");
        if (rep > 0 && ms > 0.f) {
            double tf = flop / (ms * 1e-3) / 1e12;
            if (tf > best)
                best = tf;
        }
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *out_tflops = best;
    return CVB_OK;
}
