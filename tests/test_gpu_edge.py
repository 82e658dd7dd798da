"""Edge cases of the round-2 code paths on the device: chunked host batches, lattices too small or too
thin for the batched path, NaN axis values, calls on alternating streams, top-K larger than the batch."""
import numpy as np
import pytest

from covest_b200 import workload
from covest_b200.models import BasicModel, RepeatsModel
from tests.helpers import case_hist, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rep():
    model = RepeatsModel(21, 100, case_hist(load_case('cfg2_repeats')), 0, max_error=8)
    yield model
    model.close()


def test_host_batches_larger_than_a_staging_chunk():
    """> 4 Mi points from a host array go through the staging buffers in chunks."""
    hist = case_hist(load_case('e05_basic'))
    model = BasicModel(21, 100, hist, 0, max_error=8)
    try:
        rng = np.random.default_rng(2)
        n = (1 << 22) + 12345
        pts = np.column_stack([rng.uniform(5, 20, n), rng.uniform(.01, .2, n)])
        ll = model.loglikelihood_batch(pts)
        pick = rng.choice(n, 64, replace=False)
        pick[:3] = [0, (1 << 22) - 1, n - 1]
        assert np.array_equal(ll[pick], model.loglikelihood_batch(pts[pick]))
        assert np.all(np.isfinite(ll))
    finally:
        model.close()


def test_small_and_thin_lattices_take_the_other_paths(rep):
    ctx = rep.device_context
    # fewer than 2048 points: the per-point kernel
    axes = [np.array([25., 35.]), np.array([.02, .04]), np.linspace(.4, 1, 4), np.linspace(0, 1, 4), np.linspace(.1, 1, 5)]
    a, rows = ctx.lattice_eval(axes, k_best=3)
    assert ctx.last_path_info()['kernel'] == 'cv_loglik_kernel'
    assert np.array_equal(a, ctx.loglik(workload.lattice_points(axes)))
    assert rows[0, 0] == a.max()
    # many (c, e) pairs with a single (q1, q2) per q: too thin for runs, the general plan decides
    axes = [np.geomspace(10, 90, 40), np.geomspace(.01, .09, 30), np.array([.7]), np.array([.5]), np.linspace(.1, 1, 3)]
    b, _ = ctx.lattice_eval(axes)
    assert not ctx.last_path_info()['analytic_plan']
    assert np.array_equal(b, ctx.loglik(workload.lattice_points(axes)))
    # an empty slice and a top-K larger than the batch
    none, rows = ctx.lattice_eval(axes, count=0, k_best=2)
    assert len(none) == 0 and np.all(np.isneginf(rows[:, 0]))
    few = workload.lattice_points(axes)[:5]
    rows = ctx.topk(ctx.loglik(few), few, 8)
    assert np.all(np.isfinite(rows[:5, 0])) and np.all(np.isneginf(rows[5:, 0]))


def test_nan_axis_values_behave_as_in_the_reference(rep):
    """NaN parameters pass fit_to_bounds unchanged (models.py:60-69: both comparisons are false).  A NaN
    q only matters when copies beyond the second are used: with q2 = 0 the cut-off is 2 and the value
    is finite, exactly as the reference computes it; otherwise NaN.  Nothing spreads to other points."""
    from oracle import covest_oracle as orc
    from tests.helpers import rel_err_ll
    ctx = rep.device_context
    axes = [np.geomspace(10, 90, 8), np.geomspace(.01, .09, 6), np.array([.4, np.nan, .9]), np.linspace(0, 1, 5),
            np.array([.1, .5, np.nan, 1.0])]
    ll, _ = ctx.lattice_eval(axes)
    pts = workload.lattice_points(axes)
    want = orc.Model('repeats', 21, 100, dict(rep.hist), 0, max_error=8).loglik_batch(pts, threads=8)
    assert np.array_equal(np.isnan(ll), np.isnan(want))
    assert 0 < np.isnan(want).sum() < np.isnan(pts).any(axis=1).sum()   # some NaN inputs give finite values
    assert rel_err_ll(ll, want).max() <= 1e-9
    clean = ~np.isnan(pts).any(axis=1)
    assert np.allclose(ll[clean], ctx.loglik(pts[clean]), rtol=1e-11, atol=0)   # (the subset is small: per-point kernel)


def test_calls_on_alternating_streams_are_ordered(rep):
    """Calls of one context share scratch: a call on another stream waits for the previous one."""
    import torch
    ctx = rep.device_context
    axes = [np.geomspace(10, 90, 12), np.geomspace(.01, .09, 8), np.linspace(.3, 1, 6), np.linspace(0, 1, 6),
            np.linspace(.05, 1, 8)]
    want, _ = ctx.lattice_eval(axes)
    n = len(want)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = [torch.empty(n, dtype=torch.float64, device='cuda') for _ in range(6)]
    for i, out in enumerate(outs):
        ctx.lattice_eval(axes, out_ll=out, stream=s1 if i % 2 else s2)
    torch.cuda.synchronize()
    for out in outs:
        assert np.array_equal(out.cpu().numpy(), want, equal_nan=True)


def test_long_slices_with_host_output_come_in_parts(rep):
    """cvb_lattice_eval evaluates a slice of >= 2^22 points in two to eight parts (of whole runs and whole
    (coverage, error rate) groups) when the values go to host memory, so that a part's values
    travel while the next is evaluated: the values and the best rows are those of one evaluation."""
    import torch
    ctx = rep.device_context
    axes = [np.geomspace(10, 90, 48), np.geomspace(.01, .09, 20), np.linspace(.3, 1, 10), np.linspace(0, 1, 10),
            np.linspace(.05, 1, 90)]
    total = 48 * 20 * 9000
    assert total // 2 >= 1 << 22
    dev = torch.empty(total, dtype=torch.float64, device='cuda')
    _, rows_dev = ctx.lattice_eval(axes, out_ll=dev, k_best=16)
    one = dev.cpu().numpy()
    host, rows = ctx.lattice_eval(axes, k_best=16)
    assert np.array_equal(host, one, equal_nan=True)
    assert np.array_equal(rows, rows_dev)
    # a rank's share: whole groups dealt round-robin (the bench's slicing), and a ragged tail
    block = 9000
    mine, _ = ctx.lattice_eval(axes, first=1, stride=2, block=block)
    idx = (1 + 2 * (np.arange(len(mine)) // block)) * block + np.arange(len(mine)) % block
    assert len(mine) == total // 2 and np.array_equal(mine, one[idx], equal_nan=True)
    tail, _ = ctx.lattice_eval(axes, first=3 * block, count=total - 3 * block - 1234)
    assert np.array_equal(tail, one[3 * block:total - 1234], equal_nan=True)
