#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It compiles the reference's C module
out of tree (oracle/_ref, `make -C oracle ref`), imports the reference package with three import
shims that live in a temp directory (scipy.misc.comb -> scipy.special.comb, empty matplotlib and
Bio stubs -- see SURVEY.md section 8(c)), evaluates the reference's own
BasicModel / RepeatsModel.compute_probabilities / compute_loglikelihood and
covest_poisson.truncated_poisson, and writes JSON fixtures.  Nothing from the reference is copied:
only inputs and the numbers it returned are stored.

    python tests/golden/gen_golden.py            # rewrites tests/golden/*.json, *.hist
"""
import json
import math
import multiprocessing
import os
import subprocess
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'


def import_reference():
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), 'ref'])
    shim = tempfile.mkdtemp(prefix='covest_shims_')
    os.makedirs(os.path.join(shim, 'matplotlib'))
    open(os.path.join(shim, 'matplotlib', '__init__.py'), 'w').close()
    open(os.path.join(shim, 'matplotlib', 'pyplot.py'), 'w').close()
    os.makedirs(os.path.join(shim, 'Bio'))
    with open(os.path.join(shim, 'Bio', '__init__.py'), 'w') as f:
        f.write('SeqIO = None\n')
    sys.path[:0] = [os.path.join(ROOT, 'oracle', '_ref'), shim, REF]
    warnings.simplefilter('ignore')
    import scipy.misc
    import scipy.special
    scipy.misc.comb = scipy.special.comb
    from covest import constants, models
    constants.VERBOSE = False
    import covest_poisson
    return models, covest_poisson


MODELS, POISSON = None, None


def _init():
    global MODELS, POISSON
    if MODELS is None:
        MODELS, POISSON = import_reference()


def make_model(case):
    _init()
    hist = {int(j): int(h) for j, h in case['hist']}
    cls = MODELS.RepeatsModel if case['model'] == 'repeats' else MODELS.BasicModel
    kw = dict(max_error=case['max_error'], max_cov=case.get('max_cov'))
    if case['model'] == 'repeats':
        kw['min_single_copy_ratio'] = case.get('min_q1', 0.3)
        kw['threshold'] = case.get('threshold', 1e-8)
    return cls(case['k'], case['r'], hist, case['tail'], **kw)


def _eval_point(job):
    case, point, want_probs = job
    m = make_model(case)
    ll = m.compute_loglikelihood(*point)
    probs = None
    if want_probs:
        clipped = m.fit_to_bounds(point)
        p = m.compute_probabilities(*clipped)
        probs = [p[j] for j in m.hist]
    return ll, probs


def load_hist(path):
    hist = []
    with open(path) as f:
        for line in f:
            if line.startswith('#') or not line.strip():
                continue
            a = line.split()
            hist.append([int(a[0]), int(a[1])])
    return hist


def random_points(rng, model, n, c0, lo_e=1e-3):
    pts = []
    for _ in range(n):
        c = c0 * 3 ** rng.uniform(-1, 1)
        e = math.exp(rng.uniform(math.log(lo_e), math.log(0.5)))
        if model == 'basic':
            pts.append([c, e])
        else:
            pts.append([c, e, rng.uniform(0.3, 1), rng.uniform(0, 1), rng.uniform(0.03, 1)])
    return pts


def edge_points(model, c0):
    if model == 'basic':
        return [[c0, 0.0], [c0, 0.5], [c0, 0.9], [0.001, 0.05], [c0, -0.1], [c0, 1e-12],
                [c0, 1e-6], [c0, 0.003], [c0, 0.01], [c0, 0.03], [3 * c0, 0.2], [c0 / 3, 0.4]]
    e = 0.05
    return [
        [c0, e, 1, 0, 0], [c0, e, 1, .5, .5], [c0, e, .8, .5, .5], [c0, 0, .8, .5, .5],
        [c0, .9, .1, 2, -1], [c0, .5, .3, 1, 0], [c0, e, .5, 0, .5], [c0, e, .5, 1, .5],
        [c0, e, .5, .5, 0], [c0, e, .5, .5, 1], [c0, e, .5, .5, 1e-9], [c0, e, .5, 1e-9, .5],
        [c0, e, .3, 0, 1], [c0, e, .999999999, .5, .5], [c0, 1e-12, .7, .3, .2],
        [c0, 1e-6, .7, .3, .2], [c0, 0.01, .7, .3, .05], [c0, 0.003, .6, .6, .1],
        [2 * c0, 0.02, .65, .5, .5], [c0 / 2, 0.1, .65, .5, .5], [c0, e, .65, .5, .02],
    ]


def sawtooth_points(model, k, r, n_terms=6):
    """Coverages that put o*l_0 just above a multiple of 200 (covest_poissonmodule.c:25-31)."""
    pts = []
    e = 0.01
    scale = (r - k + 1) / r * (1 - e) ** k
    for target in (200.001, 200.5, 201.0, 210.0, 400.001, 399.9999, 600.0000001)[:n_terms + 1]:
        c = target / scale
        pts.append([c, e] if model == 'basic' else [c / 3, e, .6, .5, .4])
    return pts


def synth_hist(case, theta, n_bins, n_kmers, seed, keep_zeros=False):
    """h_j ~ Poisson(N * p_j(theta)) for j = 1..n_bins, p_j from the reference itself."""
    probe = dict(case)
    probe['hist'] = [[j, 1] for j in range(1, n_bins + 1)]
    probe['tail'] = 0
    m = make_model(probe)
    p = m.compute_probabilities(*theta)
    rng = np.random.default_rng(seed)
    pj = np.array([max(p[j], 0.0) for j in range(1, n_bins + 1)])
    h = rng.poisson(n_kmers * pj)
    return [[j, int(v)] for j, v in zip(range(1, n_bins + 1), h) if keep_zeros or v > 0]


def run_case(pool, case, points, probs_at):
    jobs = [(case, p, i in probs_at) for i, p in enumerate(points)]
    res = pool.map(_eval_point, jobs, chunksize=1)
    out = dict(case)
    out['points'] = points
    out['ll'] = [r[0] for r in res]
    out['probs'] = {str(i): res[i][1] for i in sorted(probs_at)}
    return out


def main():
    _init()
    rng = np.random.default_rng(20261018)
    data = os.path.join(REF, 'tests', 'data')

    # ---- truncated_poisson known answers ---------------------------------------------------
    tp_cases = [(7.9, 5), (5, 3), (150, 150), (200, 200), (200.001, 200), (200.5, 200),
                (201, 200), (210, 200), (400.001, 400), (1000.5, 1000), (1e-9, 1), (2e-8, 1),
                (500, 480), (11356, 11356), (11400, 11400), (1e-8, 1), (1.0000001e-8, 1),
                (1e-8, 2), (3e-17, 1), (1e-300, 1), (1e-300, 2), (0.5, 300), (1e-5, 40),
                (5740.3, 5000), (5740.3, 2), (199.99999, 1), (600.00000001, 650), (37.0, 1)]
    for _ in range(400):
        lam = math.exp(rng.uniform(math.log(1e-12), math.log(9000)))
        j = int(rng.integers(1, 5001)) if rng.uniform() < 0.5 else max(1, int(lam * rng.uniform(0.5, 1.5)))
        tp_cases.append((lam, min(j, 6000)))
    tp = [[float(l), int(j), POISSON.truncated_poisson(float(l), int(j))] for l, j in tp_cases]
    with open(os.path.join(HERE, 'tp_kat.json'), 'w') as f:
        json.dump(tp, f)

    cases = []
    with multiprocessing.get_context('fork').Pool(8) as pool:
        # ---- the reference's own fixtures ---------------------------------------------------
        fixtures = {
            'e05': load_hist(os.path.join(data, 'simulated_c10_e0.05_r100_k21.hist')),
            'e05_sparse': load_hist(os.path.join(data, 'simulated_c10_e0.05_r100_k21_sparse.hist')),
            'e0': load_hist(os.path.join(data, 'simulated_c10_e0_r100_k21.hist')),
        }
        for name, hist in fixtures.items():
            for model in ('basic', 'repeats'):
                case = dict(name='%s_%s' % (name, model), model=model, k=21, r=100, max_error=8,
                            tail=0, hist=hist)
                pts = edge_points(model, 10.0) + sawtooth_points(model, 21, 100) + \
                    random_points(rng, model, 60, 10.0)
                if name == 'e05':
                    pts += [[8.774153092007053, 0.046939839227641] +
                            ([] if model == 'basic' else [.65, .5, .5]), [10, .05] +
                            ([] if model == 'basic' else [.8, .5, .5])]
                cases.append(run_case(pool, case, pts, set(range(0, len(pts), 7))))
        # all error classes (max_error=None -> S = k+1)
        case = dict(name='e05_basic_allerr', model='basic', k=21, r=100, max_error=None, tail=0,
                    hist=fixtures['e05'])
        pts = edge_points('basic', 10.0) + random_points(rng, 'basic', 30, 10.0)
        cases.append(run_case(pool, case, pts, {0, 5, 11, 20}))
        case = dict(name='e05_repeats_allerr', model='repeats', k=21, r=100, max_error=None, tail=0,
                    hist=fixtures['e05'])
        pts = edge_points('repeats', 10.0)[:8] + random_points(rng, 'repeats', 20, 10.0)
        cases.append(run_case(pool, case, pts, {2, 9}))
        # trimmed histogram with a tail (histogram.py:126-134): bins < 10 kept
        full = fixtures['e05']
        trimmed = [[j, h] for j, h in full if j < 10]
        tail = sum(h for j, h in full if j >= 10)
        for model in ('basic', 'repeats'):
            case = dict(name='e05_trim10_%s' % model, model=model, k=21, r=100, max_error=8,
                        tail=tail, hist=trimmed)
            pts = edge_points(model, 10.0)[:10] + random_points(rng, model, 40, 10.0)
            cases.append(run_case(pool, case, pts, {1, 12}))
        # max_cov bound (basic forwards it, repeats does not: models.py:23, :177)
        case = dict(name='e05_basic_maxcov', model='basic', k=21, r=100, max_error=8, tail=0,
                    hist=fixtures['e05'], max_cov=12)
        cases.append(run_case(pool, case, [[20, .05], [12, .05], [5, .05]], {0}))
        case = dict(name='e05_repeats_maxcov_minq1', model='repeats', k=21, r=100, max_error=8,
                    tail=0, hist=fixtures['e05'], max_cov=12, min_q1=0.5)
        cases.append(run_case(pool, case, [[20, .05, .4, .5, .5], [20, .05, .5, .5, .5]], {0}))

        # ---- synthetic configurations ------------------------------
        # cfg1: basic, k=21, r=100, c=10, e=.03, bins 1..300
        base = dict(model='basic', k=21, r=100, max_error=8, tail=0)
        h1 = synth_hist(base, (10, .03), 300, 1e7, 1001)
        case = dict(base, name='cfg1_basic', hist=h1)
        pts = edge_points('basic', 10.0) + random_points(rng, 'basic', 80, 10.0)
        cases.append(run_case(pool, case, pts, {0, 9, 30}))
        case = dict(base, name='cfg1_basic_dense300', hist=synth_hist(base, (10, .03), 300, 1e7, 1001, True))
        pts = random_points(rng, 'basic', 24, 10.0)
        cases.append(run_case(pool, case, pts, {3}))
        # cfg2: repeats, k=21, r=100, theta*=(30,.03,.7,.5,.5), bins 1..300
        base = dict(model='repeats', k=21, r=100, max_error=8, tail=0)
        h2 = synth_hist(base, (30, .03, .7, .5, .5), 300, 1e7, 1002)
        case = dict(base, name='cfg2_repeats', hist=h2)
        pts = edge_points('repeats', 30.0) + sawtooth_points('repeats', 21, 100) + \
            random_points(rng, 'repeats', 64, 30.0)
        cases.append(run_case(pool, case, pts, {2, 17, 40}))
        # cfg2 trimmed at 120 with tail
        h2t = [[j, h] for j, h in h2 if j < 120]
        t2 = sum(h for j, h in h2 if j >= 120)
        case = dict(base, name='cfg2_repeats_trim120', hist=h2t, tail=t2)
        pts = random_points(rng, 'repeats', 32, 30.0)
        cases.append(run_case(pool, case, pts, {5}))
        # cfg3-like: 1000 dense bins (zeros kept), a few points (each costs seconds on the CPU)
        h3 = synth_hist(base, (30, .03, .7, .5, .5), 1000, 1e8, 1003, True)
        case = dict(base, name='cfg3_repeats_dense1000', hist=h3)
        pts = [[30, .03, .7, .5, .5], [25, .05, .6, .4, .3], [40, .01, .9, .2, .8],
               [31, .03, .7, .5, .1], [90, .02, .5, .5, .5], [12, .2, .35, .9, .6],
               [30, .03, .7, .5, .05], [60, .004, .8, .1, .25]]
        cases.append(run_case(pool, case, pts, {0, 4}))
        # cfg4: k=31, r=150, theta*=(200,.01,.7,.5,.28), bins up to 5000, zeros dropped
        base = dict(model='repeats', k=31, r=150, max_error=8, tail=0)
        h4 = synth_hist(base, (200, .01, .7, .5, .28), 5000, 1e7, 1004)
        case = dict(base, name='cfg4_repeats_k31', hist=h4)
        pts = [[200, .01, .7, .5, .28], [190, .012, .65, .45, .32], [230, .008, .75, .6, .27],
               [205.3, .01, .7, .5, .5], [150, .02, .5, .3, .4], [260, .005, .8, .5, .3]]
        cases.append(run_case(pool, case, pts, {0, 3}))

    for c in cases:
        with open(os.path.join(HERE, 'loglik_%s.json' % c['name']), 'w') as f:
            json.dump(c, f)
        print(c['name'], len(c['points']), 'points', len(c['hist']), 'bins')


if __name__ == '__main__':
    main()
