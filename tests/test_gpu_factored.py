"""Parity of the factored evaluation path of the repeats model (profiles x copy weights on the FP64
tensor cores, covest_b200/csrc/factored.cu) with the reference: golden vectors of the unmodified
reference, the CPU oracle, and the per-point kernel on the same inputs.  Needs a B200."""
import math

import numpy as np
import pytest

from oracle import covest_oracle as orc
from tests.helpers import (case_ctor_kwargs, case_hist, context_for, golden_case_names, load_case,
                           rel_err_ll)

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9  # BASELINE.json north_star: per-point log-likelihood within 1e-9 relative
PATH_RTOL = 1e-11  # the two device paths differ only in summation order


def _model(case):
    return orc.Model(case['model'], case['k'], case['r'], case_hist(case), case['tail'],
                     **case_ctor_kwargs(case))


def _repeats_cases():
    return [n for n in golden_case_names() if load_case(n)['model'] == 'repeats']


KERNELS = ['gemm', 'prefix']


def _force(ctx, kernel):
    ctx.set_path({'gemm': ctx.PATH_FACTORED_GEMM, 'prefix': ctx.PATH_FACTORED_PREFIX}[kernel])


def _ran(ctx, kernel):
    info = ctx.last_path_info()
    return info['path'] == 'factored' and info['kernel'] == 'cvf_%s_kernel' % kernel


def _lattice(axes):
    return np.ascontiguousarray(np.array(np.meshgrid(*axes, indexing='ij')).reshape(len(axes), -1).T)


@pytest.mark.parametrize('kernel', KERNELS)
@pytest.mark.parametrize('name', _repeats_cases())
def test_factored_matches_reference_golden(name, kernel):
    case = load_case(name)
    m = _model(case)
    with context_for(m) as ctx:
        _force(ctx, kernel)
        got = ctx.loglik(case['points'])
        assert _ran(ctx, kernel)
    rel = rel_err_ll(got, np.array(case['ll'], dtype=float))
    assert rel.max() <= LL_RTOL, (name, int(rel.argmax()), case['points'][int(rel.argmax())],
                                  got[int(rel.argmax())], case['ll'][int(rel.argmax())])


@pytest.mark.parametrize('name,c0,n_ce,n_q', [('cfg2_repeats', 30, 12, 40), ('e05_trim10_repeats', 10, 12, 40),
                                              ('e05_repeats_allerr', 10, 6, 30),
                                              ('cfg4_repeats_k31', 200, 2, 12)])
@pytest.mark.parametrize('kernel', KERNELS)
def test_factored_grouped_points_against_oracle(name, c0, n_ce, n_q, kernel):
    """Seeded points that share (coverage, error rate) pairs, as grids and stencils do, including
    the bound and edge points of SURVEY.md section 8(d)."""
    case = load_case(name)
    m = _model(case)
    rng = np.random.default_rng(123)
    ce = np.column_stack([c0 * 3 ** rng.uniform(-1, 1, n_ce), np.exp(rng.uniform(np.log(1e-4), np.log(.5), n_ce))])
    ce[0, 1] = 0.0
    qs = np.column_stack([rng.uniform(.3, 1, n_q), rng.uniform(0, 1, n_q), rng.uniform(.02, 1, n_q)])
    qs[0] = [1.0, .5, .5]
    qs[1] = [.5, 0.0, .5]
    qs[2] = [.5, .5, 0.0]
    qs[3] = [.5, .5, 1.0]
    qs[4] = [.1, 2, -1]  # clipped to (min_q1, 1, 0)
    qs[8:, 2] = qs[5 + np.arange(n_q - 8) % 3, 2]  # three shared q: runs with many cut-offs each
    pts = np.array([[c, e, a, b, q] for c, e in ce for a, b, q in qs])
    pts = pts[rng.permutation(len(pts))]
    want = m.loglik_batch(pts, threads=8)
    with context_for(m) as ctx:
        _force(ctx, kernel)
        got = ctx.loglik(pts)
        info = ctx.last_path_info()
        assert _ran(ctx, kernel) and info['groups'] == n_ce
    inside = ~(np.isposinf(want) | np.isnan(want))  # the reference's own overflow, DESIGN.md section 3
    assert inside.sum() >= 0.8 * len(pts)
    rel = rel_err_ll(got[inside], want[inside])
    assert rel.max() <= LL_RTOL, (int(rel.argmax()), pts[inside][int(rel.argmax())])


@pytest.mark.parametrize('bins', [200, 1000])
def test_factored_equals_per_point_kernel_on_a_lattice(bins):
    """BASELINE.json configs[2] shape: the candidate lattice of initial_grid.  The automatic mode
    must choose the factored path here, and the two paths must agree far inside the tolerance."""
    case = load_case('cfg3_repeats_dense1000')
    hist = {j: h for j, h in case_hist(case).items() if j <= bins}
    m = orc.Model('repeats', case['k'], case['r'], hist, case['tail'], max_error=8)
    axes = [np.geomspace(10, 90, 6), np.geomspace(.01, .09, 4), np.linspace(.3, 1, 5),
            np.linspace(0, 1, 5), np.linspace(.05, 1, 6)]
    grid = _lattice(axes)
    with context_for(m) as ctx:
        got = ctx.loglik(grid)
        info = ctx.last_path_info()
        assert _ran(ctx, 'prefix') and info['groups'] == 24 and info['q_runs'] == 24 * 6, info
        ctx.set_path(ctx.PATH_FACTORED_GEMM)
        gemm = ctx.loglik(grid)
        assert _ran(ctx, 'gemm')
        assert rel_err_ll(got, gemm).max() <= PATH_RTOL
        ctx.set_path(ctx.PATH_PER_POINT)
        want = ctx.loglik(grid)
        assert ctx.last_path_info()['path'] == 'per-point'
        rel = rel_err_ll(got, want)
        assert rel.max() <= PATH_RTOL, (int(rel.argmax()), grid[int(rel.argmax())])
        # lattice entry point (points generated on the device) and strided slices
        ctx.set_path(ctx.PATH_AUTO)
        lat, rows = ctx.lattice_eval(axes, k_best=8)
        assert np.array_equal(lat, got, equal_nan=True)
        assert rows[0, 0] == np.nanmax(got)
        ctx.set_path(ctx.PATH_FACTORED)  # 1800 points: below the automatic mode's batch size
        part, _ = ctx.lattice_eval(axes, first=1, stride=2)
        assert ctx.last_path_info()['path'] == 'factored'
        assert np.array_equal(part, got[1::2], equal_nan=True)
    # a handful against the oracle itself
    pick = np.random.default_rng(4).choice(len(grid), 24, replace=False)
    ref = m.loglik_batch(grid[pick], threads=8)
    assert rel_err_ll(got[pick], ref).max() <= LL_RTOL


@pytest.mark.parametrize('kernel', KERNELS)
def test_factored_results_do_not_depend_on_batch_composition(kernel):
    case = load_case('cfg3_repeats_dense1000')
    m = _model(case)
    rng = np.random.default_rng(8)
    axes = [np.geomspace(12, 80, 5), np.geomspace(.01, .08, 3), np.linspace(.3, 1, 6),
            np.linspace(0, 1, 6), np.linspace(.05, 1, 8)]
    grid = _lattice(axes)
    perm = rng.permutation(len(grid))
    with context_for(m) as ctx:
        _force(ctx, kernel)
        a = ctx.loglik(grid)
        b = ctx.loglik(grid[perm])
        some = ctx.loglik(grid[100:140])
        one = ctx.loglik(grid[777:778])
    assert np.array_equal(a[perm], b, equal_nan=True)
    assert np.array_equal(some, a[100:140], equal_nan=True)
    assert one[0] == a[777]


@pytest.mark.parametrize('kernel', KERNELS)
def test_factored_edge_rows_and_small_workspace(monkeypatch, kernel):
    """NaN rows, empty copy ranges, and a profile workspace so small that the batch runs in many
    group ranges."""
    case = load_case('cfg2_repeats')
    m = _model(case)
    axes = [np.geomspace(10, 90, 8), np.geomspace(.005, .2, 4), np.linspace(.3, 1, 4),
            np.linspace(0, 1, 4), np.linspace(.02, 1, 5)]
    grid = _lattice(axes)
    grid[5, 0] = math.nan
    grid[9, 4] = math.nan
    with context_for(m) as ctx:
        ctx.set_path(ctx.PATH_PER_POINT)
        want = ctx.loglik(grid)
    monkeypatch.setenv('COVEST_B200_PROFILE_MIB', '1')
    with context_for(m) as ctx:
        _force(ctx, kernel)
        got = ctx.loglik(grid)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert math.isnan(got[5]) and math.isnan(got[9])
    assert rel_err_ll(got, want).max() <= PATH_RTOL


@pytest.mark.parametrize('kernel', KERNELS)
def test_factored_with_a_tail_term(kernel):
    """tail != 0 exercises the compensated mass sum of the epilogues (models.py:103-104)."""
    case = load_case('cfg2_repeats_trim120')
    assert case['tail'] != 0
    m = _model(case)
    axes = [np.geomspace(15, 60, 4), np.geomspace(.01, .06, 3), np.linspace(.4, .9, 4),
            np.linspace(.1, .9, 4), np.linspace(.2, .9, 4)]
    grid = _lattice(axes)
    want = m.loglik_batch(grid, threads=8)
    with context_for(m) as ctx:
        _force(ctx, kernel)
        got = ctx.loglik(grid)
    assert rel_err_ll(got, want).max() <= LL_RTOL


def test_prefix_kernel_long_runs_and_many_runs():
    """q-runs longer than a batch (512 points), more runs than a batch holds (4), tiles that end
    inside a run, and the automatic choice between the two kernels."""
    case = load_case('cfg3_repeats_dense1000')
    hist = {j: h for j, h in case_hist(case).items() if j <= 400}
    m = orc.Model('repeats', case['k'], case['r'], hist, case['tail'], max_error=8)
    rng = np.random.default_rng(77)
    rows = []
    for c, e in [(25.0, .02), (40.0, .05)]:
        for q, n in [(.5, 1500), (.3, 3), (.31, 1), (.32, 1), (.33, 1), (.34, 1), (.9, 700), (1.0, 40), (0.0, 40)]:
            for _ in range(n):
                rows.append([c, e, rng.uniform(.3, 1), rng.uniform(0, 1), q])
    pts = np.array(rows)[rng.permutation(len(rows))]
    with context_for(m) as ctx:
        ctx.set_path(ctx.PATH_FACTORED)
        got = ctx.loglik(pts)
        assert _ran(ctx, 'prefix') and ctx.last_path_info()['q_runs'] == 18
        ctx.set_path(ctx.PATH_PER_POINT)
        want = ctx.loglik(pts)
        assert rel_err_ll(got, want).max() <= PATH_RTOL
        # every point its own q: the GEMM is the better tool, and the automatic mode says so
        solo = pts.copy()
        solo[:, 4] = rng.uniform(.05, 1, len(solo))
        ctx.set_path(ctx.PATH_FACTORED)
        got = ctx.loglik(solo)
        assert _ran(ctx, 'gemm')
        _force(ctx, 'prefix')
        alt = ctx.loglik(solo)
        assert _ran(ctx, 'prefix')
        assert rel_err_ll(got, alt).max() <= PATH_RTOL
    pick = rng.choice(len(pts), 16, replace=False)
    ref = m.loglik_batch(pts[pick], threads=8)
    with context_for(m) as ctx:
        _force(ctx, 'prefix')
        sub = ctx.loglik(pts[pick])
    assert rel_err_ll(sub, ref).max() <= LL_RTOL


@pytest.mark.parametrize('seed', range(6))
def test_prefix_kernel_on_random_batch_structures(seed):
    """Random numbers of (coverage, error rate) groups, q-runs per group and points per run
    (singletons up to runs that span several batches and tiles), edge values of q1 / q2 / q mixed
    in, rows shuffled: the prefix kernel against the per-point kernel on the same rows."""
    case = load_case('cfg2_repeats' if seed % 2 else 'cfg3_repeats_dense1000')
    hist = case_hist(case)
    if seed % 2 == 0:
        hist = {j: h for j, h in hist.items() if j <= 300 + 100 * seed}
    m = orc.Model('repeats', case['k'], case['r'], hist, case['tail'] if seed % 3 else 7.0, max_error=8)
    rng = np.random.default_rng(1000 + seed)
    rows = []
    for _ in range(int(rng.integers(1, 9))):
        c, e = 30 * 3 ** rng.uniform(-1, 1), float(np.exp(rng.uniform(np.log(2e-3), np.log(.3))))
        for _ in range(int(rng.integers(1, 13))):
            q = float(rng.choice([0.0, 1.0, rng.uniform(.03, 1), rng.uniform(.2, 1)], p=[.05, .1, .35, .5]))
            n = int(rng.choice([1, 2, 3, int(rng.integers(4, 80)), int(rng.integers(80, 700)),
                                int(rng.integers(2000, 2600))], p=[.15, .1, .1, .4, .2, .05]))
            q1 = rng.choice([1.0, .3], size=n, p=[.5, .5]) * (rng.uniform(0, 1, n) < .1) + \
                rng.uniform(.3, 1, n) * 1.0
            q1 = np.clip(q1, .3, 1.0)
            q2 = np.where(rng.uniform(0, 1, n) < .1, rng.choice([0.0, 1.0], size=n), rng.uniform(0, 1, n))
            rows.append(np.column_stack([np.full(n, c), np.full(n, e), q1, q2, np.full(n, q)]))
    pts = np.vstack(rows)
    pts = pts[rng.permutation(len(pts))]
    with context_for(m) as ctx:
        _force(ctx, 'prefix')
        got = ctx.loglik(pts)
        assert _ran(ctx, 'prefix')
        ctx.set_path(ctx.PATH_PER_POINT)
        want = ctx.loglik(pts)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    rel = rel_err_ll(got, want)
    assert rel.max() <= PATH_RTOL, (seed, len(pts), int(rel.argmax()), pts[int(rel.argmax())],
                                    got[int(rel.argmax())], want[int(rel.argmax())])
