"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors produced by the
unmodified reference (tests/golden/gen_golden.py), the CPU oracle on seeded random points, and
size-independent properties at BASELINE.json's full sizes.  Needs a B200: `pytest -m gpu`."""
import math

import numpy as np
import pytest

from oracle import covest_oracle as orc
from tests.helpers import (case_ctor_kwargs, case_hist, context_for, golden_case_names, load_case,
                           rel_err_ll)

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9  # BASELINE.json north_star: per-point log-likelihood within 1e-9 relative
P_RTOL = 1e-12  # per-bin probabilities where p_j > 1e-300


def _model(case):
    return orc.Model(case['model'], case['k'], case['r'], case_hist(case), case['tail'],
                     **case_ctor_kwargs(case))


@pytest.mark.parametrize('name', golden_case_names())
def test_loglik_matches_reference_golden(name):
    case = load_case(name)
    m = _model(case)
    with context_for(m) as ctx:
        got = ctx.loglik(case['points'])
    rel = rel_err_ll(got, np.array(case['ll'], dtype=float))
    assert rel.max() <= LL_RTOL, (name, int(rel.argmax()), case['points'][int(rel.argmax())],
                                  got[int(rel.argmax())], case['ll'][int(rel.argmax())])


@pytest.mark.parametrize('name', golden_case_names())
def test_probabilities_match_reference_golden(name):
    case = load_case(name)
    m = _model(case)
    idxs = sorted(case['probs'], key=int)
    pts = [case['points'][int(i)] for i in idxs]
    with context_for(m) as ctx:
        got, ll = ctx.probs(pts, clip=True, with_loglik=True)
    for row, i in zip(got, idxs):
        want = np.array(case['probs'][i], dtype=float)
        ok = want > 1e-300
        rel = np.abs(row[ok] - want[ok]) / want[ok]
        assert rel.max() <= P_RTOL, (name, i, rel.max())
        assert np.all(row[want == 0] == 0)
    want_ll = np.array([case['ll'][int(i)] for i in idxs], dtype=float)
    assert rel_err_ll(ll, want_ll).max() <= LL_RTOL


@pytest.mark.parametrize('name,n', [('cfg2_repeats', 2000), ('cfg1_basic', 2000),
                                    ('e05_trim10_repeats', 2000), ('cfg4_repeats_k31', 48)])
def test_random_points_against_oracle(name, n):
    case = load_case(name)
    m = _model(case)
    rng = np.random.default_rng(77)
    c0 = {'cfg2_repeats': 30, 'cfg1_basic': 10, 'e05_trim10_repeats': 10, 'cfg4_repeats_k31': 200}[name]
    cols = [c0 * 3 ** rng.uniform(-1, 1, n), np.exp(rng.uniform(np.log(1e-4), np.log(.5), n))]
    if m.n_params == 5:
        cols += [rng.uniform(.3, 1, n), rng.uniform(0, 1, n), rng.uniform(.02, 1, n)]
    pts = np.column_stack(cols)
    # bound and edge points (SURVEY.md section 8(d) parity gates)
    pts[0, 1] = 0.0
    if m.n_params == 5:
        pts[1, 2] = 1.0
        pts[2, 3] = 0.0
        pts[3, 4] = 0.0
        pts[4, 4] = 1.0
        pts[5] = [c0, .9, .1, 2, -1]
    want = m.loglik_batch(pts, threads=8)
    with context_for(m) as ctx:
        got = ctx.loglik(pts)
    # Parity is defined inside the reference's own numeric domain: for rates above ~11 360 its
    # long double product overflows and it returns +inf / NaN;
    # the device path continues the formula there.
    inside = ~(np.isposinf(want) | np.isnan(want))
    assert inside.sum() >= 0.8 * n
    rel = rel_err_ll(got[inside], want[inside])
    assert rel.max() <= LL_RTOL, (int(rel.argmax()), pts[inside][int(rel.argmax())])


def test_device_buffers_and_stream_equal_host_buffers():
    import torch
    case = load_case('cfg2_repeats')
    m = _model(case)
    pts = np.array(case['points'], dtype=np.float64)
    with context_for(m) as ctx:
        host = ctx.loglik(pts)
        dev_pts = torch.from_numpy(pts).cuda()
        dev = ctx.loglik(dev_pts, stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert np.array_equal(host, dev.cpu().numpy(), equal_nan=True)


def test_results_do_not_depend_on_batch_composition():
    case = load_case('cfg3_repeats_dense1000')
    m = _model(case)
    rng = np.random.default_rng(5)
    n = 700
    pts = np.column_stack([30 * 3 ** rng.uniform(-1, 1, n), np.exp(rng.uniform(np.log(1e-3), np.log(.3), n)),
                           rng.uniform(.3, 1, n), rng.uniform(0, 1, n), rng.uniform(.05, 1, n)])
    perm = rng.permutation(n)
    with context_for(m) as ctx:
        a = ctx.loglik(pts)
        b = ctx.loglik(pts[perm])
        one = ctx.loglik(pts[17:18])
    assert np.array_equal(a[perm], b, equal_nan=True)
    assert one[0] == a[17]


def test_lattice_equals_explicit_points_and_topk():
    case = load_case('cfg2_repeats')
    m = _model(case)
    axes = [np.geomspace(10, 90, 7), np.geomspace(.005, .3, 5), np.linspace(.3, 1, 4),
            np.linspace(0, 1, 3), np.linspace(.05, 1, 4)]
    grid = np.array(np.meshgrid(*axes, indexing='ij')).reshape(5, -1).T  # last axis fastest
    with context_for(m) as ctx:
        want = ctx.loglik(grid)
        got, rows = ctx.lattice_eval(axes, k_best=16)
        assert np.array_equal(got, want, equal_nan=True)
        # strided slice, as the ranks of a multi-GPU run take it
        part, _ = ctx.lattice_eval(axes, first=3, stride=8)
        assert np.array_equal(part, want[3::8], equal_nan=True)
        # top-K: best first, ties to the lower index
        key = np.where(np.isnan(want), -np.inf, want)
        order = np.lexsort((np.arange(len(key)), -key))[:16]
        assert np.array_equal(rows[:, 0], want[order])
        assert np.array_equal(rows[:, 1:], grid[order])
        rows2 = ctx.topk(want, grid, 16)
        assert np.array_equal(rows2, rows)
        # K larger than the batch pads with (-inf, nan)
        rows3 = ctx.topk(want[:5], grid[:5], 8)
        assert np.all(np.isneginf(rows3[5:, 0])) and np.all(np.isnan(rows3[5:, 1:]))


def test_topk_of_a_large_batch_equals_the_selection_order():
    """Batches of >= 32768 points select their top-K by a radix sort (topk.cu): same total order
    as the small-batch kernel -- larger first, ties to the lower index, NaN as -inf."""
    case = load_case('e05_repeats')
    m = _model(case)
    rng = np.random.default_rng(21)
    n = 100000
    ll = -np.exp(rng.uniform(0, 20, n))
    ll[rng.choice(n, 5000, replace=False)] = -np.inf
    ll[rng.choice(n, 300, replace=False)] = np.nan
    top = np.sort(ll[np.isfinite(ll)])[-40]
    ll[rng.choice(n, 50, replace=False)] = top  # ties inside the selection
    pts = rng.uniform(0.1, 1, (n, 5))
    key = np.where(np.isnan(ll), -np.inf, ll)
    order = np.lexsort((np.arange(n), -key))
    with context_for(m) as ctx:
        rows = ctx.topk(ll, pts, 128)
        assert np.array_equal(rows[:, 0], key[order[:128]])
        assert np.array_equal(rows[:, 1:], pts[order[:128]])
        few = ctx.topk(ll[:1000], pts[:1000], 16)  # small batch: the selection kernel
        o2 = np.lexsort((np.arange(1000), -key[:1000]))[:16]
        assert np.array_equal(few[:, 1:], pts[o2])
        allbad = np.full(40000, np.nan)
        rows = ctx.topk(allbad, pts[:40000], 4)
        assert np.all(np.isneginf(rows[:, 0])) and np.array_equal(rows[:, 1:], pts[:4])
        # the radix selection at its largest K, the radix sort beyond it, and a constant batch
        # (every key equal: the index bits of the 96-bit keys decide)
        for k in (1024, 2000):
            rows = ctx.topk(ll, pts, k)
            assert np.array_equal(rows[:, 0], key[order[:k]]) and np.array_equal(rows[:, 1:], pts[order[:k]])
        same = np.full(70000, -3.5)
        rows = ctx.topk(same, pts[:70000], 100)
        assert np.all(rows[:, 0] == -3.5) and np.array_equal(rows[:, 1:], pts[:100])


def test_full_size_properties_cfg3():
    """BASELINE.json configs[2] shape (repeats, 1000 dense bins) at a batch the oracle could not
    finish: q1 = 1 makes the repeats model the basic model (SURVEY.md section 8(c) invariants),
    clipping is idempotent, and every value is a negative number or -inf (a point under which an
    observed bin has probability zero), never NaN or +inf."""
    case = load_case('cfg3_repeats_dense1000')
    m = _model(case)
    basic = orc.Model('basic', case['k'], case['r'], case_hist(case), case['tail'], max_error=8)
    rng = np.random.default_rng(9)
    n = 20000
    pts = np.column_stack([30 * 3 ** rng.uniform(-1, 1, n), np.exp(rng.uniform(np.log(1e-3), np.log(.3), n)),
                           rng.uniform(.3, 1, n), rng.uniform(0, 1, n), rng.uniform(.05, 1, n)])
    with context_for(m) as ctx, context_for(basic) as bctx:
        ll = ctx.loglik(pts)
        assert not np.any(np.isnan(ll)) and np.all(ll < 0) and np.isfinite(ll).mean() > 0.5
        q1one = pts.copy()
        q1one[:, 2] = 1.0
        a = ctx.loglik(q1one[:4000])
        b = bctx.loglik(q1one[:4000, :2])
        assert rel_err_ll(a, b).max() <= 1e-13
        outside = pts[:4000].copy()
        outside[:, 1] += 1.0   # error rate above its bound 0.5
        outside[:, 4] -= 2.0   # q below 0
        clipped = outside.copy()
        clipped[:, 1] = 0.5
        clipped[:, 4] = 0.0
        assert np.array_equal(ctx.loglik(outside), ctx.loglik(clipped))
        # probabilities are non-negative (their sum may exceed 1: the reference's denominator
        # quirk for rates above 200 is reproduced) and the likelihood follows from them
        p, ll2 = ctx.probs(pts[:64], clip=True, with_loglik=True)
        assert np.array_equal(ll2, ll[:64])
        assert np.all(p >= 0) and not np.any(np.isnan(p))
        h = np.array([v for v in case_hist(case).values()], dtype=float)
        with np.errstate(divide='ignore'):
            manual = np.array([np.sum(h[h > 0] * np.log(row[h > 0])) for row in p])
        assert rel_err_ll(manual, ll2).max() <= 1e-12


def test_error_reporting():
    from covest_b200.engine import DeviceError, LikelihoodContext
    with pytest.raises(DeviceError):
        LikelihoodContext(0, 21, 100, 8, [], [], 0, None, ((.01, None), (0, .5)), [1.0] * 8)
    with pytest.raises(DeviceError):
        LikelihoodContext(0, 21, 100, 8, [1, 1], [2, 3], 0, None, ((.01, None), (0, .5)), [1.0] * 8)
    case = load_case('e05_basic')
    with context_for(_model(case)) as ctx:
        assert len(ctx.loglik(np.zeros((0, 2)))) == 0
        assert math.isnan(ctx.loglik([[math.nan, .05]])[0])


def test_device_merge_of_rank_blocks_equals_the_host_formulation():
    """cvb_merge_rows (the merge after the all-gather of a multi-GPU round) against
    parallel.merge_topk's torch formulation on CPU tensors: ties, NaN, -inf padding."""
    import torch

    from covest_b200 import parallel
    rng = np.random.default_rng(21)
    for n, k in ((128, 64), (512, 64), (100, 7), (5, 8)):
        rows = rng.normal(size=(n, 6))
        rows[:, 0] = np.round(rows[:, 0], 1)             # many ties on the log-likelihood
        rows[rng.choice(n, n // 8, replace=False), 0] = -np.inf
        rows[rng.choice(n, max(1, n // 16), replace=False), 0] = np.nan
        rows[rng.choice(n, max(1, n // 16), replace=False), 3] = np.nan
        rows[n // 2:n // 2 + n // 4] = rows[:n // 4]      # duplicated rows (the same point on two ranks)
        want = parallel.merge_topk(torch.from_numpy(rows), k).numpy()
        got = parallel.merge_topk(torch.from_numpy(rows).cuda(), k).cpu().numpy()
        assert got.shape == want.shape
        assert np.array_equal(got, want, equal_nan=True), (n, k)
