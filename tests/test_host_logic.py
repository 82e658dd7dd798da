"""Host-side logic around the hot path, on CPU: the reference-interface mirror (models registry,
bounds, histogram pre-processing, candidate generators, estimator drivers, report) against golden
values produced by the unmodified reference (tests/golden/gen_host_golden.py).

The device is replaced, in these tests only, by the CPU oracle plugged in at the one seam the
product evaluates through (`loglikelihood_batch`); the product code itself never does that."""
import json
import os
import random

import numpy as np
import pytest
import yaml

from covest_b200 import constants, data, grid, histogram, models
from covest_b200.covest import CoverageEstimator, LaunchBatcher, build_parser
from oracle import covest_oracle as orc
from tests.helpers import GOLDEN

constants.VERBOSE = False

with open(os.path.join(GOLDEN, 'host_golden.json')) as f:
    HOST = json.load(f)


def fixture(name):
    return {int(j): int(h) for j, h in HOST['fixtures'][name]}


class OracleBacked:
    """Mixin: route the batched evaluators to the CPU oracle (tests only)."""

    def _oracle(self):
        return orc.Model(self._kind, self.k, self.r, self.hist, self.tail, max_error=self.max_error,
                         max_cov=self.bounds[0][1], threshold=self.threshold if self.repeats else 1e-8,
                         min_single_copy_ratio=self.bounds[2][0] if self.repeats else 0.3)

    def loglikelihood_batch(self, points):
        pts = np.asarray(points, dtype=np.float64).reshape(-1, self.param_count)
        return self._oracle().loglik_batch(pts, threads=4)


# same class names as the product's, so that short_name() reports 'basic' / 'repeats'
OBasic = type('BasicModel', (OracleBacked, models.BasicModel), {})
ORepeats = type('RepeatsModel', (OracleBacked, models.RepeatsModel), {})


# ---- models: registry and bounds (reference tests/test_models.py) ----------------------------
def test_select_model_full_prefix_invalid():
    assert models.select_model('basic') is models.BasicModel
    assert models.select_model('repeats') is models.RepeatsModel
    assert models.select_model('r') is models.RepeatsModel
    assert models.select_model('repeat') is models.RepeatsModel
    assert models.select_model('b') is models.BasicModel
    with pytest.raises(ValueError):
        models.select_model('x')
    assert set(models.models) == {'basic', 'repeats'}


def test_model_attributes_match_reference_conventions():
    b = models.BasicModel(21, 100, {1: 5, 2: 3}, 0, max_error=8, max_cov=50)
    assert b.params == ('coverage', 'error_rate') and b.param_count == 2
    assert b.bounds == ((0.01, 50), (0, 0.5)) and b.defaults == (1, 0.25)
    assert b.max_error == 8 and len(b.comb) == 22 and not b.repeats
    assert models.BasicModel(21, 100, {1: 1}, 0).max_error == 22
    r = models.RepeatsModel(21, 100, {1: 5, 2: 3}, 0, max_error=8, max_cov=50, min_single_copy_ratio=0.4)
    assert r.params == ('coverage', 'error_rate', 'q1', 'q2', 'q') and r.repeats
    assert r.bounds == ((0.01, None), (0, 0.5), (0.4, 1), (0, 1), (0, 1))  # max_cov is not forwarded
    assert r.defaults == (1, 0.25, 0.7, 0.5, 0.5) and r.threshold == 1e-8
    assert r.short_name() == 'repeats' and b.short_name() == 'basic'
    assert r.fit_to_bounds([10, .9, .1, 2, -1]) == [10, 0.5, 0.4, 1, 0]
    assert r.check_bounds([10, .1, .5, .5, .5]) and not r.check_bounds([10, .6, .5, .5, .5])
    assert b.correct_c(10) == 10 * 80 / 100
    bo = r.get_b_o(0.5, 0.4, 0.3)
    assert bo(0) == 0 and bo(1) == 0.5 and bo(2) == 0.5 * 0.4 and bo(5) == 0.5 * 0.6 * 0.3 * 0.7 ** 2
    assert r.get_hist_threshold(bo, 1e-8) == 2  # max(hist) = 2 caps the cut-off
    assert [float(v) for v in b.comb[:3]] == [1.0, 63.0, 1890.0]


# ---- histogram pre-processing ------------------------------------------------------------------
@pytest.mark.parametrize('name', ['e05', 'e05_sparse', 'e0'])
def test_histogram_preprocessing_matches_reference(name):
    want = HOST['histogram'][name]
    hist = fixture(name)
    assert list(histogram.compute_coverage_apx(hist, 21, 100)) == want['apx']
    assert histogram.get_trim(hist) == want['trim']
    assert histogram.get_trim(hist, True) == want['trim_ignore_last']
    h, tail, sf, c, e = histogram.process_histogram(hist, 21, 100, sample_factor=1)
    assert [[j, v] for j, v in h.items()] == want['processed']['hist']
    assert (tail, sf, [c, e]) == (want['processed']['tail'], want['processed']['sample_factor'],
                                  want['processed']['guess'])
    h3, tail3 = histogram.trim_hist(hist, 10)
    assert [[j, v] for j, v in h3.items()] == want['trim10']['hist'] and tail3 == want['trim10']['tail']


@pytest.mark.parametrize('name', ['cfg1_basic', 'cfg2_repeats'])
def test_process_histogram_on_synthetic_configs(name):
    with open(os.path.join(GOLDEN, 'loglik_%s.json' % name)) as f:
        case = json.load(f)
    hist = {int(j): int(h) for j, h in case['hist']}
    want = HOST['histogram'][name]
    h, tail, sf, c, e = histogram.process_histogram(hist, 21, 100, sample_factor=1)
    assert [[j, v] for j, v in h.items()] == want['processed']['hist']
    assert (tail, [c, e]) == (want['processed']['tail'], want['processed']['guess'])
    h, tail, sf, c, e = histogram.process_histogram(hist, 21, 100, sample_factor=1, trim=0)
    assert (len(h), tail, [c, e]) == (want['trim_t0']['n_bins'], want['trim_t0']['tail'], want['trim_t0']['guess'])


def test_compute_coverage_apx_error_free_histogram():
    # the inline case of the reference's tests/test_histograms.py:17-46
    hist = dict(zip(range(1, 25), [2909, 10891, 28824, 56698, 92099, 122998, 137748, 137507, 124723,
                                   100866, 72467, 47639, 29893, 17119, 9026, 4713, 2077, 767, 288,
                                   139, 49, 27, 37, 16]))
    c, e = histogram.compute_coverage_apx(hist, 21, 100)
    assert abs(c - 10) < 1 and abs(e) < 0.01
    assert histogram.compute_coverage_apx({}, 21, 100) == (0.0, 1.0)


def test_sample_histogram_halves_coverage_and_trim_conserves_counts():
    random.seed(4)
    hist = fixture('e0')
    half = histogram.sample_histogram(hist, 2)
    c0, _ = histogram.compute_coverage_apx(hist, 21, 100)
    c1, _ = histogram.compute_coverage_apx(half, 21, 100)
    assert abs(c1 - c0 / 2) < 0.5
    trimmed, tail = histogram.trim_hist(hist, 12)
    assert sum(trimmed.values()) + tail == sum(hist.values())
    assert histogram.trim_hist(hist, 1000) == (hist, 0)
    pd = histogram.poisson_dist(3.0, 5)
    assert abs(pd[2] - 0.22404180765538775) < 1e-15 and histogram.poisson_dist(0, 3) == [0.0] * 3


# ---- data ------------------------------------------------------------------------------------
def test_histogram_file_round_trip(tmp_path):
    hist = fixture('e05_sparse')
    fn = str(tmp_path / 'a.hist')
    data.save_histogram(hist, fn, {'tool': 'x', 'sample_factor': 3})
    got, meta = data.load_histogram(fn)
    assert got == hist and meta == {'tool': 'x', 'sample_factor': '3'}
    data.save_histogram(hist, fn)
    assert data.load_histogram(fn) == (hist, {})
    (tmp_path / 'bad.hist').write_text('1 2\nfoo bar\n')
    with pytest.raises(data.InvalidFormatException):
        data.load_histogram(str(tmp_path / 'bad.hist'))
    (tmp_path / 'empty.hist').write_text('')
    assert data.load_histogram(str(tmp_path / 'empty.hist')) == ({}, {})


def test_report_matches_reference_report():
    est = HOST['estimator']['basic']
    hist = {int(j): int(h) for j, h in est['hist']}
    model = OBasic(21, 100, hist, est['tail'], max_error=8)
    rep = data.print_output(fixture('e05'), model, est['success'], 1, est['x'], est['guess'], [None, None],
                            silent=True)
    want = HOST['estimator']['basic_report']
    assert set(rep) == set(want)
    for key, value in want.items():
        if isinstance(value, float):
            assert rep[key] == pytest.approx(value, rel=1e-12), key
        else:
            assert rep[key] == value, key
    assert yaml.safe_load(yaml.dump(rep))['genome_size'] == want['genome_size']


# ---- candidate generators --------------------------------------------------------------------
def test_initial_grid_matches_reference_with_the_same_seed():
    for key in ('initial_grid', 'initial_grid_fix'):
        g = HOST[key]
        random.seed(g['seed'])
        bounds = [tuple(b) for b in g['bounds']]
        pts = grid.initial_grid(g['guess'], count=g['count'], bounds=bounds, fix=g.get('fix'))
        assert [list(p) for p in pts] == g['points']
    assert grid.initial_grid([1, 2], count=0) == []


def test_optimize_grid_matches_reference_on_a_toy_objective():
    g = HOST['optimize_grid_toy']

    def toy(x):
        return (x[0] - 7.3) ** 2 + 40 * (x[1] - 0.031) ** 2 + 0.5 * (x[0] - 7.3) * (x[1] - 0.031)

    calls = []

    class Batched:
        def __call__(self, x):
            return toy(x)

        def batch(self, pts):
            calls.append(len(pts))
            return [toy(p) for p in pts]

    bounds = [tuple(b) for b in g['bounds']]
    assert list(grid.optimize_grid(toy, g['start'], bounds=bounds)) == g['result']
    assert list(grid.optimize_grid(Batched(), g['start'], bounds=bounds)) == g['result']
    assert max(calls) == 36 and calls[0] == 1  # one launch per round: 6^2 candidates
    cands = grid.grid_candidates([10.0, 0.0, 0.5], 1.1, 3, bounds=((0.01, 12), (0, .5), (0, 1)), fix=[None, None, 0.7])
    assert len(cands) == 4 * 6 and all(c[1] == 0.0 and c[2] == 0.7 for c in cands)  # 12.1 and 13.31 exceed 12


# ---- estimator drivers -----------------------------------------------------------------------
@pytest.mark.parametrize('name', ['basic', 'repeats'])
def test_single_start_follows_the_reference_optimizer(name):
    """Batching the finite-difference stencil must not change L-BFGS-B's path: same point and
    evaluation count as the stock reference run (which evaluates one point at a time)."""
    est = HOST['estimator'][name]
    hist = {int(j): int(h) for j, h in est['hist']}
    cls = OBasic if name == 'basic' else ORepeats
    model = cls(21, 100, hist, est['tail'], max_error=8)
    ce = CoverageEstimator(model)
    r = ce._optimize(est['guess'])
    assert [float(v) for v in r.x] == est['x']
    assert float(r.fun) == est['fun'] and r.nfev == est['nfev'] and r.nit == est['nit']
    assert ce.launches < est['nfev']  # stencils were batched


def test_multi_start_threads_equal_sequential_runs():
    est = HOST['estimator']['basic']
    hist = {int(j): int(h) for j, h in est['hist']}
    model = OBasic(21, 100, hist, est['tail'], max_error=8)
    ce = CoverageEstimator(model)
    random.seed(7)
    starts = grid.initial_grid(est['guess'], count=5, bounds=ce.bounds)
    seq = [ce._optimize(s) for s in starts]
    before = ce.launches
    par = ce._optimize_many(starts)
    merged = ce.launches - before
    for a, b in zip(seq, par):
        assert np.array_equal(a.x, b.x) and a.fun == b.fun and a.nfev == b.nfev
    assert merged < sum(2 * r.nit for r in seq)  # launches were shared between the starts
    random.seed(7)
    x, ok = ce.compute_coverage(est['guess'], starting_points=5)
    best = min(seq, key=lambda r: r.fun)
    assert list(x) == list(best.x) and ok == best.success


def test_compute_coverage_with_grid_matches_reference():
    est = HOST['estimator']['basic']
    hist = {int(j): int(h) for j, h in est['hist']}
    model = OBasic(21, 100, hist, est['tail'], max_error=8)
    x, ok = CoverageEstimator(model).compute_coverage(est['guess'], starting_points=1, use_grid_search=True)
    assert [float(v) for v in x] == HOST['estimator']['basic_grid']['x']
    assert ok == HOST['estimator']['basic_grid']['success']


def test_err_scale_and_fix_overlay():
    hist = fixture('e05')
    model = OBasic(21, 100, hist, 0, max_error=8)
    ce = CoverageEstimator(model, err_scale=10, fix=[None, 0.04])
    assert ce.bounds[1] == (0, 5.0)
    got = ce.likelihood_f([10.0, 0.3])
    want = -model.loglikelihood_batch([[10.0, 0.004]])[0]  # fixed value, then divided by err_scale
    assert got == want
    assert list(ce.likelihood_f.batch([[10.0, 0.3], [9.0, 0.1]])) == [want, -model.loglikelihood_batch([[9.0, 0.004]])[0]]


def test_polish_reaches_the_well_determined_optimum():
    """the optimum of the basic objective on the reference's fixture,
    found by Newton polish from different starts (stock L-BFGS-B stops ~4e-5 away)."""
    hist = fixture('e05')
    model = OBasic(21, 100, hist, 0, max_error=8)
    ce = CoverageEstimator(model)
    for start in ([10.018173967731428, 0.04999051411001868], [9.5, 0.045], [10.4, 0.052]):
        x, f = ce.polish(start)
        assert x[0] == pytest.approx(10.018595075, rel=2e-9)
        assert x[1] == pytest.approx(0.0499913066, rel=2e-8)
        assert f == pytest.approx(3678682.5783989, rel=1e-12)


def test_launch_batcher_propagates_errors_and_retirement():
    import threading

    def bad(points):
        raise RuntimeError('boom')

    b = LaunchBatcher(bad, 2)
    errs = []

    def client():
        try:
            b.submit([[1.0]])
        except RuntimeError as exc:
            errs.append(str(exc))

    ts = [threading.Thread(target=client) for _ in range(2)]
    [t.start() for t in ts]
    [t.join(5) for t in ts]
    assert errs == ['boom', 'boom']
    ok = LaunchBatcher(lambda pts: [p[0] * 2 for p in pts], 2)
    ok.retire()  # the other client never submits
    assert ok.submit([[1.0], [2.0]]) == [2.0, 4.0] and ok.launches == 1


def test_cli_flags_are_the_reference_flags():
    p = build_parser()
    a = p.parse_args(['x.hist', '-m', 'repeat', '-k', '31', '-r', '150', '-sp', '16', '-t', '0', '-sf', '1', '-g',
                      '-c', '30', '-e', '.03', '-p', '.7', '.5', '-f', '-ll', '-so', '-es', '2', '-mq1', '.2',
                      '-M', '300', '-rs', '1000', '-T', '3'])
    assert (a.model, a.kmer_size, a.read_length, a.starting_points, a.trim, a.sample_factor, a.grid) == \
        ('repeat', 31, 150, 16, 0, 1, True)
    assert (a.coverage, a.error_rate, list(a.params), a.fix, a.ll_only, a.start_original) == \
        (30.0, .03, [.7, .5], True, True, True)
    assert (a.error_scale, a.min_q1, a.max_coverage, a.reads_size, a.thread_count) == (2.0, .2, 300, 1000, 3)
    d = p.parse_args(['x.hist'])
    assert (d.model, d.kmer_size, d.read_length, d.starting_points, d.grid, d.trim, d.sample_factor) == \
        ('basic', 21, 100, 1, False, None, None)


def test_workload_flop_accounting_of_the_factored_paths():
    """bench.py's roofline numerators on a lattice small enough to count by
    hand: copies per point from the cut-off, groups by (c, e), q-runs by (c, e, q)."""
    from covest_b200 import workload
    from covest_b200.models import RepeatsModel
    hist = {j: 10 for j in range(1, 101)}
    model = RepeatsModel(21, 100, hist, 0, max_error=8)
    axes = [np.array([10.0, 20.0]), np.array([0.02]), np.array([0.5, 1.0]), np.array([0.0, 0.5]),
            np.array([0.2, 0.6])]
    pts = workload.lattice_points(axes)
    assert pts.shape == (16, 5)
    copies = np.maximum(workload.copy_cutoff(pts, max(hist), model.threshold) - 1, 0)
    # q1 = 1 or q2 = 0 with ... : the cut-off follows models.py:185-191
    assert copies[(pts[:, 2] == 1.0)].max() == 1            # b(2) = 0 <= threshold: only copy 1
    w = workload.factored_flop(model, pts, n_bins=100, counted_bins=100)
    assert w['groups'] == 2 and w['q_runs'] == 4
    assert w['gemm_flop'] == float(np.sum(2.0 * 100 * copies + 64.0 * 100))
    per_run_max = [copies[(pts[:, 0] == c) & (pts[:, 4] == q)].max() for c in (10.0, 20.0) for q in (0.2, 0.6)]
    assert w['prefix_flop'] == 16 * (6.0 * 100 + 64.0 * 100) + 2.0 * 100 * sum(max(m - 2, 0) for m in per_run_max)
    assert w['profile_flop'] == 2.0 * 8 * 100 * sum(copies[pts[:, 0] == c].max() for c in (10.0, 20.0))
    model.close()


def test_grid_rounds_as_arrays_keep_the_order_of_itertools_product():
    import itertools
    from covest_b200 import grid
    axes = grid.grid_axes([10.0, 0.05, 0.5, 0.0, 0.9], 1.1, 3, bounds=((0.01, None), (0, .5), (.3, 1), (0, 1), (0, 1)))
    assert [len(a) for a in axes] == [6, 6, 6, 6, 4]   # 0.0 stays 0.0 six times; 0.9 * 1.1^d cut at 1
    rows = grid._product_rows(axes)
    assert rows.shape == (6 * 6 * 6 * 6 * 4, 5)
    assert [tuple(r) for r in rows[:50]] == list(itertools.product(*axes))[:50]
    assert tuple(rows[-1]) == list(itertools.product(*axes))[-1]
    assert grid._product_rows([[1.0], [], [2.0]]).shape == (0, 3)
    assert grid.grid_candidates([10.0, 0.05], 1.1, 1) == list(itertools.product(*grid.grid_axes([10.0, 0.05], 1.1, 1)))


# ---- lock-step multi-start, device grid rounds (additions) -----------------------------------
def test_lockstep_optimizer_reaches_the_polished_optimum_from_every_start():
    """c = 10.018595075, e = 0.0499913066, f = 3678682.5783989 on the
    reference's fixture; all starts advance per launch."""
    from covest_b200.optimizer import lockstep_minimize
    hist = fixture('e05')
    ce = CoverageEstimator(OBasic(21, 100, hist, 0, max_error=8))
    random.seed(7)
    starts = grid.initial_grid([8.77, 0.0469], count=8, bounds=ce.bounds)
    results, launches = lockstep_minimize(ce.likelihood_batch, starts, ce.bounds)
    assert launches <= 2 * 15 + 1
    for r in results:
        assert r.success
        assert r.x[0] == pytest.approx(10.018595075, rel=2e-7) and r.x[1] == pytest.approx(0.0499913066, rel=2e-7)
        assert r.fun == pytest.approx(3678682.5783989, rel=1e-12)
    # a fixed coordinate stays, bounds hold
    results, _ = lockstep_minimize(ce.likelihood_batch, [[9.0, 0.04]], ce.bounds, fixed=[False, True])
    assert results[0].x[1] == 0.04 and results[0].fun > 3678682.58
    results, _ = lockstep_minimize(ce.likelihood_batch, [[9.0, 0.04]], [(0.01, 9.5), (0, .5)])
    assert results[0].x[0] == 9.5


def test_refine_starts_and_lockstep_compute_coverage():
    hist = fixture('e05')
    model = ORepeats(21, 100, hist, 0, max_error=8)
    ce = CoverageEstimator(model, optimizer='lockstep')
    random.seed(3)
    x, ok = ce.compute_coverage([8.77, 0.0469, .65, .5, .5], starting_points=4)
    assert ok and ce.launches < 200
    # at least as good as the stock single-start run of the reference
    assert ce.likelihood_f(x) <= 3678677.5264701946 + 1e-3
    xs, fun, ok2, table = ce.refine_starts([[10.0, .05, .9, .5, .5], [9.0, .04, .5, .5, .5]])
    assert table.shape == (2, 7) and fun == table[:, 0].min() and list(xs) == list(table[np.argmin(table[:, 0]), 2:])


def test_grid_round_through_lattice_best_equals_the_sequential_bookkeeping():
    """optimize_grid with an objective that offers lattice_best (the device path: only the best
    candidate of a round comes back) walks the same centres as the reference's bookkeeping."""
    g = HOST['optimize_grid_toy']

    def toy(x):
        return (x[0] - 7.3) ** 2 + 40 * (x[1] - 0.031) ** 2 + 0.5 * (x[0] - 7.3) * (x[1] - 0.031)

    rounds = []

    class OnDevice:
        def __call__(self, x):
            return toy(x)

        def batch(self, pts):
            return [toy(p) for p in pts]

        def lattice_best(self, axes):
            import itertools
            cands = list(itertools.product(*axes))
            vals = [toy(c) for c in cands]
            i = int(np.argmin(vals))  # first minimum
            rounds.append(len(cands))
            return vals[i], tuple(cands[i])

    bounds = [tuple(b) for b in g['bounds']]
    assert list(grid.optimize_grid(OnDevice(), g['start'], bounds=bounds)) == g['result']
    assert len(rounds) >= 17 and max(rounds) == 36
