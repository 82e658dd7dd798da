"""The row layout of the batched path (csrc/factored.h, CvfSlots): profile rows keep the 64-bin lines
of the histogram that hold a count and one half line with the sums of all other lines.  Bins
without counts enter the result through the mass only (models.py:103-104), so this has to hold
exactly where the mass matters: a histogram with a tail whose mass sits in bins without counts."""
import numpy as np
import pytest

from covest_b200 import workload
from covest_b200.models import RepeatsModel
from tests.helpers import rel_err_ll

pytestmark = pytest.mark.gpu

K, R, BINS, TAIL = 21, 100, 700, 5000


def sparse_hist():
    """700 keys; counts in 1..90 and a few just above 200: lines 0, 1 and 3 of the 11 with bins hold counts."""
    rng = np.random.default_rng(77)
    hist = {j: 0 for j in range(1, BINS + 1)}
    for j in range(1, 91):
        hist[j] = int(rng.integers(1, 10 ** 6 // j))
    for j in (200, 201, 204):
        hist[j] = int(rng.integers(1, 50))
    return hist


def axes():
    # coverage 8 keeps the mass in the bins with counts, 25 and 33 move much of it to bins without (and
    # 330 copies at coverage 33 stay below the rates where the reference overflows, ~11 360)
    return [np.array([8., 25., 33.]), np.array([.01, .05]), np.array([.35, .6, .85, 1.0]),
            np.array([0., .3, .7, 1.]), np.array([.05, .2, .45, .7, .9, 1.])]


def oracle_values(hist, tail, pts):
    from oracle import covest_oracle as orc
    return orc.Model('repeats', K, R, dict(hist), tail, max_error=8).loglik_batch(pts, threads=8)


@pytest.mark.parametrize('tail', [TAIL, 0])
def test_rows_without_the_empty_lines_match_the_oracle(tail):
    hist = sparse_hist()
    model = RepeatsModel(K, R, hist, tail, max_error=8)
    try:
        ctx = model.device_context
        pts = workload.lattice_points(axes())
        want = oracle_values(hist, tail, pts)
        # |ll| is ~1e8 here and the tail term at most 5000 * 36: its conditioning (tests/bigpoints.py) is far
        # inside the plain gate
        for path in (ctx.PATH_FACTORED_PREFIX, ctx.PATH_FACTORED_GEMM, ctx.PATH_PER_POINT):
            ctx.set_path(path)
            got, _ = ctx.lattice_eval(axes())
            info = ctx.last_path_info()
            if path != ctx.PATH_PER_POINT:
                assert info['path'] == 'factored' and info['row_slots'] == 3 * 64 + 32   # 3 lines with counts + the sums
            assert np.array_equal(np.isinf(got), np.isinf(want))
            assert rel_err_ll(got, want).max() <= 1e-9, (path, float(rel_err_ll(got, want).max()))
    finally:
        model.close()


def test_the_mass_of_the_summed_lines_is_the_mass_of_their_bins(monkeypatch):
    """Same histogram, rows with every line (COVEST_B200_ROWS=full): the tail term sees the same mass.
    At coverage 33 most of the mass is in bins without counts; 1 - mass is what the tail term takes the
    logarithm of, so the two layouts agree only if the sums carry the mass to the last bits."""
    hist = sparse_hist()
    pts = workload.lattice_points(axes())
    compact = RepeatsModel(K, R, hist, TAIL, max_error=8)
    try:
        ctx = compact.device_context
        ctx.set_path(ctx.PATH_FACTORED_PREFIX)
        a = ctx.loglik(pts)
        slots = ctx.last_path_info()['row_slots']
        no_tail = RepeatsModel(K, R, hist, 0, max_error=8)
        try:
            c0 = no_tail.device_context
            c0.set_path(c0.PATH_FACTORED_PREFIX)
            tail_term = a - c0.loglik(pts)   # tail * log(1 - mass)
        finally:
            no_tail.close()
    finally:
        compact.close()
    monkeypatch.setenv('COVEST_B200_ROWS', 'full')
    full = RepeatsModel(K, R, hist, TAIL, max_error=8)
    try:
        ctx = full.device_context
        ctx.set_path(ctx.PATH_FACTORED_PREFIX)
        b = ctx.loglik(pts)
        assert ctx.last_path_info()['row_slots'] == 1024 and slots == 3 * 64 + 32   # 700 bins: one block of 16 lines
    finally:
        full.close()
    fin = np.isfinite(a)
    assert np.array_equal(fin, np.isfinite(b))
    assert rel_err_ll(a, b).max() <= 1e-12
    assert np.abs(tail_term[fin]).max() > 1e4   # the tail term is there and large


def _lined_hist(lines, bins):
    """Counts in every bin of the first `lines` 64-bin lines, zeros up to `bins`."""
    rng = np.random.default_rng(1000 + lines)
    return {j: (int(rng.integers(1, 5000)) if j <= 64 * lines else 0) for j in range(1, bins + 1)}


@pytest.mark.parametrize('lines,bins,tail', [(1, 1000, 0), (2, 1000, 7), (2, 128, 0), (3, 1000, 0), (4, 1000, 11), (5, 1000, 0),
                                             (6, 1000, 0), (7, 1000, 5), (8, 1000, 0), (9, 1000, 0), (11, 1000, 3),
                                             (12, 1000, 0), (13, 1000, 0), (16, 1024, 0), (20, 2000, 9)])
def test_every_geometry_of_the_prefix_kernel(lines, bins, tail):
    """The prefix kernel picks warps per CTA and slots per thread by the length of a profile row
    (factored.cu, cvf_eval): 3, 5, 4, 7, 9, ... half-lines here.  Every geometry gives the values of the
    GEMM path and of the per-point kernel (which is checked against the oracle elsewhere)."""
    hist = _lined_hist(lines, bins)
    model = RepeatsModel(K, R, hist, tail, max_error=8)
    try:
        ctx = model.device_context
        # coverage grows with the bins that have counts; at most ~60 copies of at most 150 x 0.8 per copy
        # stay below the rates where the reference overflows (~11 360)
        ax = [np.array([min(12. * lines, 150.), min(25. * lines, 150.)]), np.array([.02, .06]), np.array([.4, .7, 1.0]),
              np.array([0., .5, 1.]), np.linspace(.3, 1, 14)]
        got = {}
        for path in (ctx.PATH_FACTORED_PREFIX, ctx.PATH_FACTORED_GEMM, ctx.PATH_PER_POINT):
            ctx.set_path(path)
            got[path], _ = ctx.lattice_eval(ax)
            if path == ctx.PATH_FACTORED_PREFIX:
                info = ctx.last_path_info()
                assert info['kernel'] == 'cvf_prefix_kernel'
                table_lines = -(-bins // 64) if bins > 128 else 2
                table_lines = 16 * -(-table_lines // 16) if bins > 128 else table_lines
                assert info['row_slots'] == (64 * lines + 32 if lines < table_lines else 64 * lines)
        want = got[ctx.PATH_PER_POINT]
        assert np.isfinite(want).sum() >= 30   # (one or two copies cannot reach the far bins: -inf there)
        for path in (ctx.PATH_FACTORED_PREFIX, ctx.PATH_FACTORED_GEMM):
            assert np.array_equal(np.isfinite(got[path]), np.isfinite(want))
            assert rel_err_ll(got[path], want).max() <= 1e-11, (path, float(rel_err_ll(got[path], want).max()))
    finally:
        model.close()
