"""BASELINE.json configs[2] at its full size -- the benchmark's own batch: the 40 x 25 x 10 x 10 x 10
lattice of 10^6 (coverage, error rate, q1, q2, q) points over 1000 dense bins -- through
properties that do not need the oracle to finish 10^6 evaluations (it does 40 per second):
agreement of the three device paths, the q1 = 1 identity with the basic model, invariance under
permutation, the top-K contract, and a seeded sample against the oracle itself.  Needs a B200."""
import numpy as np
import pytest

from covest_b200 import workload
from covest_b200.models import BasicModel, RepeatsModel
from oracle import covest_oracle as orc
from tests.helpers import rel_err_ll

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9     # BASELINE.json north_star
PATH_RTOL = 1e-11  # device paths against each other


@pytest.fixture(scope='module')
def cfg3():
    cfg = workload.CONFIGS['cfg3']
    hist = workload.synthetic_histogram('cfg3')
    model = RepeatsModel(cfg['k'], cfg['r'], hist, 0, max_error=8)
    axes = workload.lattice_axes(cfg['theta'], n_c=40, n_e=25)
    pts = workload.lattice_points(axes)
    assert pts.shape == (1000000, 5) and len(hist) == 1000
    ctx = model.device_context
    ll = ctx.loglik(pts)   # host buffers: copies and evaluation pipelined in slices
    yield dict(cfg=cfg, hist=hist, model=model, axes=axes, pts=pts, ll=ll)
    model.close()


def test_the_lattice_takes_the_prefix_kernel(cfg3):
    import torch
    ctx, ll = cfg3['model'].device_context, cfg3['ll']
    dev = ctx.loglik(torch.from_numpy(cfg3['pts']).cuda())   # device buffers: one evaluation
    info = ctx.last_path_info()
    assert info['path'] == 'factored' and info['kernel'] == 'cvf_prefix_kernel', info
    assert info['groups'] == 40 * 25 and info['q_runs'] == 40 * 25 * 10
    assert np.array_equal(dev.cpu().numpy(), ll)
    assert not np.any(np.isnan(ll)) and not np.any(np.isposinf(ll)) and np.all(ll < 0)


def test_the_three_device_paths_agree(cfg3):
    ctx, pts, ll = cfg3['model'].device_context, cfg3['pts'], cfg3['ll']
    try:
        ctx.set_path(ctx.PATH_FACTORED_GEMM)
        gemm = ctx.loglik(pts)
        assert ctx.last_path_info()['kernel'] == 'cvf_gemm_kernel'
        ctx.set_path(ctx.PATH_PER_POINT)
        direct = ctx.loglik(pts)
        assert ctx.last_path_info()['path'] == 'per-point'
    finally:
        ctx.set_path(ctx.PATH_AUTO)
    assert np.array_equal(np.isfinite(ll), np.isfinite(gemm))
    assert rel_err_ll(ll, gemm).max() <= PATH_RTOL
    # the per-point kernel sums all terms of a bin at once: where a bin with a count has a
    # probability in the subnormal range (far-off points, ll below twice the best one) one unit of
    # the subnormal grid is ln 2 in the sum, or the difference between a tiny number and zero
    # (DESIGN.md section 3.2); a few dozen of the 10^6 points
    rel = rel_err_ll(ll, direct)
    off = ~(rel <= PATH_RTOL)
    assert off.mean() <= 5e-4, off.mean()
    assert np.all(np.minimum(ll[off], direct[off]) < 1.9 * ll.max())
    both = off & np.isfinite(ll) & np.isfinite(direct)
    assert rel[both].max(initial=0.0) <= 1e-8


def test_q1_equal_one_is_the_basic_model(cfg3):
    """b(1) = 1 and every other copy weight 0 (models.py:193-208): the 10^5 lattice points with
    q1 = 1 must reproduce the basic model at their (coverage, error rate)."""
    cfg, pts, ll = cfg3['cfg'], cfg3['pts'], cfg3['ll']
    sel = pts[:, 2] == 1.0
    assert sel.sum() == 100000
    basic = BasicModel(cfg['k'], cfg['r'], cfg3['hist'], 0, max_error=8)
    try:
        ce = np.unique(pts[sel][:, :2], axis=0)
        want = dict(zip(map(tuple, ce), basic.device_context.loglik(ce)))
    finally:
        basic.close()
    ref = np.array([want[tuple(r)] for r in pts[sel][:, :2]])
    assert rel_err_ll(ll[sel], ref).max() <= 1e-13


def test_values_do_not_depend_on_the_order_or_the_rest_of_the_batch(cfg3):
    ctx, pts, ll = cfg3['model'].device_context, cfg3['pts'], cfg3['ll']
    perm = np.random.default_rng(3).permutation(len(pts))
    assert np.array_equal(ctx.loglik(pts[perm]), ll[perm])
    part = ctx.loglik(pts[250000:500000])
    assert ctx.last_path_info()['kernel'] == 'cvf_prefix_kernel'
    assert np.array_equal(part, ll[250000:500000])
    lat, rows = ctx.lattice_eval(cfg3['axes'], k_best=64)
    assert np.array_equal(lat, ll)
    order = np.lexsort((np.arange(len(ll)), -ll))[:64]   # descending, ties to the lower index
    assert np.array_equal(rows[:, 0], ll[order]) and np.array_equal(rows[:, 1:], pts[order])


def test_a_seeded_sample_against_the_oracle(cfg3):
    cfg, pts, ll = cfg3['cfg'], cfg3['pts'], cfg3['ll']
    m = orc.Model('repeats', cfg['k'], cfg['r'], {int(j): int(v) for j, v in cfg3['hist'].items()}, 0,
                  max_error=8)
    rng = np.random.default_rng(11)
    # stay where the copy cut-off is moderate: one oracle evaluation costs O(copies x bins^2)
    cheap = np.nonzero(pts[:, 4] >= 0.3)[0]
    pick = np.concatenate([rng.choice(cheap, 40, replace=False), [int(np.argmax(ll))]])
    want = m.loglik_batch(pts[pick], threads=8)
    inside = np.isfinite(want)
    assert inside.sum() >= 30
    assert rel_err_ll(ll[pick][inside], want[inside]).max() <= LL_RTOL


def test_cfg5_shape_two_full_passes():
    """BASELINE.json configs[4] shape on one rank, reduced in points: 2000 dense bins (two full
    passes of the prefix kernel over the bins), 20 q-runs per (coverage, error rate)."""
    cfg = workload.CONFIGS['cfg5']
    hist = workload.synthetic_histogram('cfg5')
    assert len(hist) == 2000
    model = RepeatsModel(cfg['k'], cfg['r'], hist, 0, max_error=8)
    axes = [np.geomspace(15, 60, 4), np.geomspace(.01, .08, 3), np.linspace(.3, 1, 5), np.linspace(0, 1, 5),
            np.linspace(.05, 1, 20)]
    pts = workload.lattice_points(axes)
    try:
        ctx = model.device_context
        ll = ctx.loglik(pts)
        info = ctx.last_path_info()
        assert info['kernel'] == 'cvf_prefix_kernel' and info['groups'] == 12 and info['q_runs'] == 240, info
        ctx.set_path(ctx.PATH_PER_POINT)
        direct = ctx.loglik(pts)
        ctx.set_path(ctx.PATH_FACTORED_GEMM)
        gemm = ctx.loglik(pts)
    finally:
        model.close()
    assert rel_err_ll(ll, gemm).max() <= PATH_RTOL
    assert rel_err_ll(ll, direct).max() <= PATH_RTOL
    m = orc.Model('repeats', cfg['k'], cfg['r'], {int(j): int(v) for j, v in hist.items()}, 0, max_error=8)
    pick = np.nonzero(pts[:, 4] >= 0.4)[0][::97][:12]
    want = m.loglik_batch(pts[pick], threads=8)
    assert rel_err_ll(ll[pick], want).max() <= LL_RTOL
