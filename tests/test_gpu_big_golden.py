"""SURVEY.md section 8(d) parity gate: >= 10^4 seeded points per BASELINE.json config -- random
points of the initial_grid box including the small-q region, points of the benchmark lattices,
rates just above 200 n, bound and edge points (tests/bigpoints.py) -- against values of the pinned
C oracle computed offline (tests/golden/big_<cfg>.npz, gen_big_golden.py), through the C ABI on all
three device paths, at 1e-9 relative with no carve-out inside the reference's numeric domain
(points where the reference itself overflows to +inf / NaN, rates above ~11 360, are outside it)."""
import numpy as np
import pytest

from tests import bigpoints
from tests.helpers import rel_err_ll

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9  # BASELINE.json north_star


def _model(big):
    from covest_b200.models import BasicModel, RepeatsModel
    cfg = big['cfg']
    cls = RepeatsModel if cfg['model'] == 'repeats' else BasicModel
    return cls(cfg['k'], cfg['r'], big['hist'], 0, max_error=8)


def check(name, path):
    big = bigpoints.load_big(name)
    model = _model(big)
    try:
        ctx = model.device_context
        ctx.set_path(path)
        got = ctx.loglik(big['points'])
        info = ctx.last_path_info()
    finally:
        model.close()
    want = big['ll']
    inside = ~(np.isposinf(want) | np.isnan(want))
    assert inside.mean() >= 0.9, inside.mean()
    rel = rel_err_ll(got[inside], want[inside])
    bad = np.nonzero(~(rel <= LL_RTOL))[0]
    worst = int(np.argmax(rel))
    assert len(bad) == 0, (name, info['kernel'], len(bad), big['points'][inside][worst].tolist(),
                           float(got[inside][worst]), float(want[inside][worst]))
    return info


@pytest.mark.parametrize('name', ['cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg5'])
def test_big_golden_default_path(name):
    check(name, 0)


@pytest.mark.parametrize('name', ['cfg2', 'cfg3', 'cfg4', 'cfg5'])
def test_big_golden_per_point_kernel(name):
    assert check(name, 1)['kernel'] == 'cv_loglik_kernel'


@pytest.mark.parametrize('name', ['cfg2', 'cfg3', 'cfg4', 'cfg5'])
def test_big_golden_factored_gemm(name):
    assert check(name, 3)['kernel'] == 'cvf_gemm_kernel'


@pytest.mark.parametrize('name', ['cfg2', 'cfg3', 'cfg4', 'cfg5'])
def test_big_golden_factored_prefix(name):
    assert check(name, 4)['kernel'] == 'cvf_prefix_kernel'
