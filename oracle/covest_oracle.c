/*
 * covest_oracle.c -- CPU restatement of CovEst's likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (covest_b200/, the C-ABI
 * library) may link, import or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 * tests/golden/*.json, which were produced by importing the unmodified
 * reference package from /root/reference together with its C module compiled
 * from c_src/covest_poissonmodule.c (script: tests/golden/gen_golden.py), and
 * -- when oracle/_ref/ holds that compiled module -- against the module itself.
 *
 * Each function cites the reference lines (relative to /root/reference) whose
 * arithmetic it follows.  The arithmetic is deliberately bug-compatible: x87
 * long double products, the staged e^200 division, the `1.0 - exp(-l)`
 * cancellation, non-renormalised copy-number weights, left-to-right sums.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CVO_MAX_PARAMS 5
#define CVO_STAGE 200 /* constants.py:1 / covest_poissonmodule.c:5 */

/* ------------------------------------------------------------------ */
/* c_src/covest_poissonmodule.c:7-35  truncated_poisson(l, j)          */
/* The zero-truncated Poisson pmf as the reference evaluates it: the   */
/* numerator is a long double running product of DOUBLE quotients l/i, */
/* the denominator is e^l - 1 with l first brought under 200 by        */
/* dividing the numerator by e^200 per step; tiny reduced l keeps the  */
/* ORIGINAL l as denominator.  l == 0 or NaN is meant to give 0        */
/* (c:15-17 passes an int through varargs; the value is only ever      */
/* multiplied by a zero weight, so 0 is the intended reading).         */
/* ------------------------------------------------------------------ */
double cvo_truncated_poisson(double rate, int j)
{
    if (rate == 0 || rate != rate)
        return 0.0;
    long double numer = 1;
    long double denom = rate;
    for (int i = 1; i <= j; i++)
        numer *= rate / i; /* double division, long double product (c:22-24) */
    while (rate > CVO_STAGE && numer > 0) { /* c:25-28 */
        numer /= expl(CVO_STAGE);
        rate -= CVO_STAGE;
    }
    if (rate > 1e-8 && numer > 0) /* c:29-31 */
        denom = expl(rate) - 1;
    return (double)(numer / denom);
}

/* Same values, bit for bit, for j = 1..max_j in one pass: the running   */
/* product of c:22-24 for j is the product for j-1 times (l / j), in the */
/* same order, so extending it one factor at a time reproduces every     */
/* per-call result exactly.  out[j-1] = truncated_poisson(rate, j).      */
static void tp_ladder(double rate, int max_j, double *out)
{
    if (rate == 0 || rate != rate) {
        for (int j = 1; j <= max_j; j++)
            out[j - 1] = 0.0;
        return;
    }
    long double numer = 1;
    for (int j = 1; j <= max_j; j++) {
        numer *= rate / j;
        long double staged = numer;
        double reduced = rate;
        long double denom = rate;
        while (reduced > CVO_STAGE && staged > 0) {
            staged /= expl(CVO_STAGE);
            reduced -= CVO_STAGE;
        }
        if (reduced > 1e-8 && staged > 0)
            denom = expl(reduced) - 1;
        out[j - 1] = (double)(staged / denom);
    }
}

/* ------------------------------------------------------------------ */
/* Python's math.fsum (models.py:103): exactly rounded sum (Shewchuk). */
/* ------------------------------------------------------------------ */
static double exact_sum(const double *v, int n)
{
    double stackbuf[64];
    double *part = stackbuf;
    int cap = 64, np = 0;
    double inf_sum = 0.0;
    int saw_nan = 0;
    for (int t = 0; t < n; t++) {
        double x = v[t];
        if (x != x) {
            saw_nan = 1;
            continue;
        }
        if (isinf(x)) {
            inf_sum += x;
            continue;
        }
        int keep = 0;
        for (int q = 0; q < np; q++) {
            double y = part[q];
            if (fabs(x) < fabs(y)) {
                double tmp = x;
                x = y;
                y = tmp;
            }
            volatile double hi = x + y;
            volatile double yr = hi - x;
            double lo = y - yr;
            if (lo != 0.0)
                part[keep++] = lo;
            x = hi;
        }
        np = keep;
        if (x != 0.0) {
            if (np == cap) {
                cap *= 2;
                double *grown = (double *)malloc(sizeof(double) * cap);
                memcpy(grown, part, sizeof(double) * np);
                if (part != stackbuf)
                    free(part);
                part = grown;
            }
            part[np++] = x;
        }
    }
    double result;
    if (saw_nan || inf_sum != inf_sum) {
        result = NAN;
    } else if (inf_sum != 0.0) {
        result = inf_sum;
    } else {
        /* sum the non-overlapping partials from the top, then fix the     */
        /* half-way case as CPython's math_fsum does                        */
        double hi = 0.0;
        int q = np;
        if (q > 0) {
            hi = part[--q];
            double lo = 0.0;
            while (q > 0) {
                double x = hi;
                double y = part[--q];
                volatile double s = x + y;
                volatile double yr = s - x;
                hi = s;
                lo = y - yr;
                if (lo != 0.0)
                    break;
            }
            if (q > 0 && ((lo < 0.0 && part[q - 1] < 0.0) || (lo > 0.0 && part[q - 1] > 0.0))) {
                double y = lo * 2.0;
                volatile double x = hi + y;
                volatile double yr = x - hi;
                if (y == yr)
                    hi = x;
            }
        }
        result = hi;
    }
    if (part != stackbuf)
        free(part);
    return result;
}

/* ------------------------------------------------------------------ */
/* Python's builtin sum(), as the interpreter that ran the reference   */
/* for the golden vectors evaluates it (CPython 3.12.3, scipy 1.18).   */
/* Two behaviours occur on the path:                                   */
/*  - models.py:88-97, :224-241: the summands are numpy.float64        */
/*    scalars (scipy's comb() returns one and it propagates), which    */
/*    are not exact `float` objects, so sum() takes its generic path:  */
/*    plain left-to-right addition;                                    */
/*  - models.py:105-107: the summands `h * safe_log(p)` are exact      */
/*    Python floats, which CPython >= 3.12 adds with Neumaier's        */
/*    compensated algorithm, applying the compensation at the end      */
/*    (older interpreters -- the reference's CI ran 3.4-3.6 -- added   */
/*    left to right; the two differ by ~1e-16 relative).               */
/* ------------------------------------------------------------------ */
static int g_compensated_sum = 1;
void cvo_set_compensated_sum(int on) { g_compensated_sum = on; }

typedef struct {
    double total, comp;
    int count, compensated;
} pysum;

static void pysum_init(pysum *s, int exact_floats)
{
    s->total = 0.0;
    s->comp = 0.0;
    s->count = 0;
    s->compensated = exact_floats && g_compensated_sum;
}

static void pysum_add(pysum *s, double x)
{
    if (s->count++ == 0) { /* int 0 + first float */
        s->total = 0.0 + x;
        return;
    }
    if (!s->compensated) {
        s->total += x;
        return;
    }
    volatile double t = s->total + x;
    if (fabs(s->total) >= fabs(x))
        s->comp += (s->total - t) + x;
    else
        s->comp += (x - t) + s->total;
    s->total = t;
}

static double pysum_value(const pysum *s)
{
    if (s->comp != 0.0 && isfinite(s->comp))
        return s->total + s->comp;
    return s->total;
}

/* utils.py:32-35 safe_log */
static double safe_log(double x)
{
    if (x <= 0)
        return -INFINITY;
    return log(x);
}

/* ------------------------------------------------------------------ */
/* The model description the Python classes carry (models.py:19-31,    */
/* :175-183): k, r, the bins of `hist` in dict order with their counts,*/
/* tail, comb[s] = C(k,s)*3^s as the reference computed it, max_error, */
/* bounds (NaN = open end), and for the repeats model the threshold.   */
/* ------------------------------------------------------------------ */
typedef struct {
    int model_kind; /* 0 = basic (models.py:17), 1 = repeats (models.py:173) */
    int k, r;
    int n_err;      /* max_error, models.py:28-31 */
    int n_bins;
    const int *bin_j;
    const double *bin_h;
    double tail;
    const double *comb; /* n_err entries, models.py:25 */
    double threshold;   /* models.py:183; NaN = None */
    double lo[CVO_MAX_PARAMS], hi[CVO_MAX_PARAMS];
} cvo_model;

static int n_params(const cvo_model *m) { return m->model_kind ? 5 : 2; }

/* models.py:60-69 fit_to_bounds */
static void clip_to_bounds(const cvo_model *m, const double *in, double *out)
{
    int np = n_params(m);
    for (int i = 0; i < np; i++) {
        double v = in[i];
        if (m->lo[i] == m->lo[i] && v < m->lo[i])
            v = m->lo[i];
        else if (m->hi[i] == m->hi[i] && v > m->hi[i])
            v = m->hi[i];
        out[i] = v;
    }
}

/* models.py:71-79: c_k = c*(r-k+1)/r ; l_s = c_k * 3**-s * (1-err)**(k-s) * err**s */
static void error_class_rates(const cvo_model *m, double c, double err, double *l_s)
{
    double ck = c * (double)(m->r - m->k + 1) / (double)m->r;
    for (int s = 0; s < m->n_err; s++) {
        double third = (s == 0) ? 1.0 : pow(3.0, (double)-s);
        double keep = pow(1.0 - err, (double)(m->k - s));
        double miss = pow(err, (double)s);
        l_s[s] = ck * third * keep * miss;
    }
}

static int max_bin(const cvo_model *m)
{
    int mx = 0;
    for (int b = 0; b < m->n_bins; b++)
        if (m->bin_j[b] > mx)
            mx = m->bin_j[b];
    return mx;
}

/* models.py:193-208 get_b_o */
static double copy_weight(int o, double q1, double q2, double q)
{
    double two = (1 - q1) * q2;
    double many = (1 - q1) * (1 - q2) * q;
    if (o == 0)
        return 0;
    if (o == 1)
        return q1;
    if (o == 2)
        return two;
    return many * pow(1 - q, (double)(o - 3));
}

/* models.py:185-191 get_hist_threshold */
static int copy_cutoff(const cvo_model *m, double q1, double q2, double q)
{
    int top = max_bin(m);
    if (m->threshold == m->threshold) {
        for (int o = 1; o < top; o++)
            if (copy_weight(o, q1, q2, q) <= m->threshold)
                return o;
    }
    return top;
}

/* Per-bin mixture probabilities.  mode 0 calls truncated_poisson once per  */
/* (bin, copy, error class) exactly as models.py:92-97 / :235-241 do; mode 1 */
/* uses tp_ladder (bit-identical, linear in max bin).  out_p[b] pairs with  */
/* bin_j[b].                                                                 */
static void mixture_probs(const cvo_model *m, const double *par, int mode, double *out_p)
{
    int S = m->n_err;
    int B = m->n_bins;
    double l_s[64];
    error_class_rates(m, par[0], par[1], l_s);

    int first_o = 1, end_o = 2; /* basic: a single "copy" with weight 1 */
    double q1 = 1, q2 = 0, q = 0;
    if (m->model_kind) {
        q1 = par[2];
        q2 = par[3];
        q = par[4];
        end_o = copy_cutoff(m, q1, q2, q);
    }
    int n_o = end_o - first_o;
    if (n_o < 0)
        n_o = 0;
    int top = max_bin(m);

    double *ladder = NULL;
    if (mode == 1)
        ladder = (double *)malloc(sizeof(double) * (size_t)(top > 0 ? top : 1) * S);
    double *per_copy = (double *)malloc(sizeof(double) * (size_t)(n_o > 0 ? n_o : 1) * B);
    double *a = (double *)malloc(sizeof(double) * S);
    double *rate = (double *)malloc(sizeof(double) * S);

    for (int oi = 0; oi < n_o; oi++) {
        int o = first_o + oi;
        /* models.py:87-90 (basic) / :220-232 (repeats) */
        pysum tot;
        pysum_init(&tot, 0);
        for (int s = 0; s < S; s++) {
            double arg = m->model_kind ? (double)o * -l_s[s] : -l_s[s];
            a[s] = m->comb[s] * (1.0 - exp(arg));
            pysum_add(&tot, a[s]);
            rate[s] = m->model_kind ? (double)o * l_s[s] : l_s[s];
        }
        double total = pysum_value(&tot);
        if (total == 0)
            total = 1; /* utils.py:25-29 fix_zero */
        for (int s = 0; s < S; s++)
            a[s] = a[s] / total;
        if (mode == 1)
            for (int s = 0; s < S; s++)
                tp_ladder(rate[s], top, ladder + (size_t)s * top);
        for (int b = 0; b < B; b++) {
            int j = m->bin_j[b];
            pysum inner;
            pysum_init(&inner, 0);
            for (int s = 0; s < S; s++) {
                double tp;
                if (mode == 1)
                    tp = (j >= 1) ? ladder[(size_t)s * top + (j - 1)] : cvo_truncated_poisson(rate[s], j);
                else
                    tp = cvo_truncated_poisson(rate[s], j);
                pysum_add(&inner, a[s] * tp);
            }
            per_copy[(size_t)oi * B + b] = pysum_value(&inner);
        }
    }
    for (int b = 0; b < B; b++) {
        pysum acc;
        pysum_init(&acc, 0);
        for (int oi = 0; oi < n_o; oi++) {
            double term = per_copy[(size_t)oi * B + b];
            if (m->model_kind)
                term = copy_weight(first_o + oi, q1, q2, q) * term; /* models.py:236 */
            pysum_add(&acc, term);
        }
        out_p[b] = pysum_value(&acc);
    }
    free(rate);
    free(a);
    free(per_copy);
    free(ladder);
}

/* models.py:100-107 compute_loglikelihood */
static double loglik_one(const cvo_model *m, const double *par_in, int mode, double *scratch)
{
    double par[CVO_MAX_PARAMS];
    clip_to_bounds(m, par_in, par);
    mixture_probs(m, par, mode, scratch);
    double mass = exact_sum(scratch, m->n_bins);
    if (!(mass < 1)) /* min(1, x) keeps the 1 unless x < 1 (also for NaN) */
        mass = 1;
    double tail_term = 0;
    if (mass < 1)
        tail_term = m->tail * safe_log(1 - mass);
    pysum acc;
    pysum_init(&acc, 1);
    for (int b = 0; b < m->n_bins; b++) {
        double h = m->bin_h[b];
        if (h == 0)
            continue;
        pysum_add(&acc, h * safe_log(scratch[b]));
    }
    return pysum_value(&acc) + tail_term;
}

/* ------------------------------------------------------------------ */
/* exported entry points (ctypes, see oracle/covest_oracle.py)         */
/* ------------------------------------------------------------------ */
static void fill_model(cvo_model *m, int model_kind, int k, int r, int n_err, int n_bins,
                       const int *bin_j, const double *bin_h, double tail, const double *comb,
                       double threshold, const double *bounds)
{
    m->model_kind = model_kind;
    m->k = k;
    m->r = r;
    m->n_err = n_err;
    m->n_bins = n_bins;
    m->bin_j = bin_j;
    m->bin_h = bin_h;
    m->tail = tail;
    m->comb = comb;
    m->threshold = threshold;
    for (int i = 0; i < CVO_MAX_PARAMS; i++) {
        m->lo[i] = NAN;
        m->hi[i] = NAN;
    }
    for (int i = 0; i < n_params(m); i++) {
        m->lo[i] = bounds[2 * i];
        m->hi[i] = bounds[2 * i + 1];
    }
}

/* per-bin probabilities at ONE parameter point (no clipping, like the
 * reference's compute_probabilities) */
int cvo_probs(int model_kind, int k, int r, int n_err, int n_bins, const int *bin_j,
              const double *comb, double threshold, const double *par, int mode, double *out_p)
{
    cvo_model m;
    double open_bounds[2 * CVO_MAX_PARAMS];
    for (int i = 0; i < 2 * CVO_MAX_PARAMS; i++)
        open_bounds[i] = NAN;
    if (n_err > 64)
        return -1;
    fill_model(&m, model_kind, k, r, n_err, n_bins, bin_j, NULL, 0.0, comb, threshold, open_bounds);
    mixture_probs(&m, par, mode, out_p);
    return 0;
}

typedef struct {
    const cvo_model *m;
    const double *par;
    double *out;
    long n_points;
    int mode;
    long *next;
    pthread_mutex_t *lock;
} worker_arg;

static void *worker(void *p)
{
    worker_arg *w = (worker_arg *)p;
    int np = n_params(w->m);
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(w->m->n_bins > 0 ? w->m->n_bins : 1));
    for (;;) {
        pthread_mutex_lock(w->lock);
        long i = (*w->next)++;
        pthread_mutex_unlock(w->lock);
        if (i >= w->n_points)
            break;
        w->out[i] = loglik_one(w->m, w->par + (size_t)i * np, w->mode, scratch);
    }
    free(scratch);
    return NULL;
}

/* log-likelihood of n_points parameter rows (row-major, 2 or 5 columns),
 * evaluated on n_threads host threads. */
int cvo_loglik_batch(int model_kind, int k, int r, int n_err, int n_bins, const int *bin_j,
                     const double *bin_h, double tail, const double *comb, double threshold,
                     const double *bounds, long n_points, const double *par, int mode,
                     int n_threads, double *out_ll)
{
    cvo_model m;
    if (n_err > 64)
        return -1;
    fill_model(&m, model_kind, k, r, n_err, n_bins, bin_j, bin_h, tail, comb, threshold, bounds);
    if (n_threads < 1)
        n_threads = 1;
    long next = 0;
    pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
    worker_arg arg = {&m, par, out_ll, n_points, mode, &next, &lock};
    if (n_threads == 1) {
        worker(&arg);
        return 0;
    }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 0; t < n_threads; t++)
        pthread_create(&tid[t], NULL, worker, &arg);
    for (int t = 0; t < n_threads; t++)
        pthread_join(tid[t], NULL);
    free(tid);
    return 0;
}
