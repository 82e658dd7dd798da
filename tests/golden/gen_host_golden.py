#!/usr/bin/env python
"""Golden vectors for the HOST logic around the hot path, produced by the unmodified reference
(run in the build container only; needs /root/reference): histogram pre-processing, candidate
generators, optimiser drivers, stock CLI reports.  Writes tests/golden/host_golden.json.

    python tests/golden/gen_host_golden.py
"""
import contextlib
import io
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden  # noqa: E402  (reuses the import shims)


def toy_objective(x):
    """A smooth picklable objective for the reference's optimize_grid (it pickles `fn`)."""
    return (x[0] - 7.3) ** 2 + 40 * (x[1] - 0.031) ** 2 + 0.5 * (x[0] - 7.3) * (x[1] - 0.031)


def main():
    gen_golden._init()
    from covest import constants, grid, histogram
    from covest import covest as ref_covest
    from covest import data as ref_data
    constants.VERBOSE = False
    data = os.path.join(gen_golden.REF, 'tests', 'data')
    out = {}
    fixtures = {name: dict(gen_golden.load_hist(os.path.join(data, fn))) for name, fn in (
        ('e05', 'simulated_c10_e0.05_r100_k21.hist'),
        ('e05_sparse', 'simulated_c10_e0.05_r100_k21_sparse.hist'),
        ('e0', 'simulated_c10_e0_r100_k21.hist'))}
    out['fixtures'] = {k: [[j, h] for j, h in v.items()] for k, v in fixtures.items()}

    # histogram pre-processing (deterministic parts)
    hp = {}
    for name, hist in fixtures.items():
        hist = {int(j): int(h) for j, h in hist.items()}
        c, e = histogram.compute_coverage_apx(hist, 21, 100)
        h2, tail, sf, gc, ge = histogram.process_histogram(hist, 21, 100, sample_factor=1)
        h3, tail3 = histogram.trim_hist(hist, 10)
        hp[name] = dict(apx=[c, e], trim=histogram.get_trim(hist), trim_ignore_last=histogram.get_trim(hist, True),
                        processed=dict(hist=[[j, h] for j, h in h2.items()], tail=tail, sample_factor=sf,
                                       guess=[gc, ge]),
                        trim10=dict(hist=[[j, h] for j, h in h3.items()], tail=tail3))
    # cfg1/cfg2 synthetic histograms through process_histogram with -sf 1
    for name in ('cfg1_basic', 'cfg2_repeats'):
        with open(os.path.join(HERE, 'loglik_%s.json' % name)) as f:
            case = json.load(f)
        hist = {int(j): int(h) for j, h in case['hist']}
        h2, tail, sf, gc, ge = histogram.process_histogram(hist, 21, 100, sample_factor=1)
        hp[name] = dict(processed=dict(hist=[[j, h] for j, h in h2.items()], tail=tail, sample_factor=sf,
                                       guess=[gc, ge]), trim_t0=None)
        h4, tail4, sf4, gc4, ge4 = histogram.process_histogram(hist, 21, 100, sample_factor=1, trim=0)
        hp[name]['trim_t0'] = dict(n_bins=len(h4), tail=tail4, guess=[gc4, ge4])
    out['histogram'] = hp

    # candidate generators
    random.seed(12345)
    out['initial_grid'] = dict(seed=12345, guess=[8.7, 0.047, 0.65, 0.5, 0.5], count=6,
                               bounds=[[0.01, None], [0, 0.5], [0.3, 1], [0, 1], [0, 1]],
                               points=grid.initial_grid([8.7, 0.047, 0.65, 0.5, 0.5], count=6, bounds=(
                                   (0.01, None), (0, 0.5), (0.3, 1), (0, 1), (0, 1))))
    random.seed(99)
    out['initial_grid_fix'] = dict(seed=99, guess=[10.0, 0.05], count=4, bounds=[[0.01, None], [0, 0.5]],
                                   fix=[None, 0.05],
                                   points=grid.initial_grid([10.0, 0.05], count=4, bounds=((0.01, None), (0, 0.5)),
                                                            fix=[None, 0.05]))
    res = grid.optimize_grid(toy_objective, [5.0, 0.05], bounds=((0.01, None), (0, 0.5)), n_threads=2)
    out['optimize_grid_toy'] = dict(start=[5.0, 0.05], bounds=[[0.01, None], [0, 0.5]], result=list(res))

    # the reference's estimator on its fixture: stock L-BFGS-B, basic and repeats, and the grid
    est = {}
    for model_name in ('basic', 'repeats'):
        hist = {int(j): int(h) for j, h in fixtures['e05'].items()}
        h2, tail, sf, gc, ge = histogram.process_histogram(hist, 21, 100)
        cls = gen_golden.MODELS.models[model_name]
        model = cls(21, 100, h2, tail, max_error=8)
        guess = list(model.defaults)
        guess[:2] = gc, ge
        ce = ref_covest.CoverageEstimator(model)
        r = ce._optimize(guess)
        est[model_name] = dict(guess=guess, x=[float(v) for v in r.x], fun=float(r.fun), nfev=int(r.nfev),
                               nit=int(r.nit), success=bool(r.success), tail=tail,
                               hist=[[j, h] for j, h in h2.items()])
        if model_name == 'basic':
            g = ce.compute_coverage(guess, starting_points=1, use_grid_search=True, n_threads=4)
            est['basic_grid'] = dict(x=[float(v) for v in g[0]], success=bool(g[1]))
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                rep = ref_data.print_output(hist, model, bool(r.success), sf, [float(v) for v in r.x], guess,
                                            [None, None], silent=True)
            est['basic_report'] = {k: (v if not hasattr(v, 'item') else v.item()) for k, v in rep.items()}
    out['estimator'] = est

    with open(os.path.join(HERE, 'host_golden.json'), 'w') as f:
        json.dump(out, f)
    print('wrote host_golden.json')


if __name__ == '__main__':
    main()
