"""CPU checks of the kernel's arithmetic, without a GPU.

tests/host_math/emulate.cpp runs the very phase functions the sm_100a kernel runs
(covest_b200/csrc/cvpoint.h) in a serial loop over the thread index.  These tests compare that
against the golden vectors of the reference and against the oracle -- they validate the
*formulation* (log-domain seeds, scaled recurrences, tables).  The CUDA path itself is checked by
the `-m gpu` tests through the C-ABI.
"""
import math

import numpy as np
import pytest

from oracle import covest_oracle as orc
from tests import emulation as emu
from tests.helpers import (case_ctor_kwargs, case_hist, golden_case_names, load_case,
                           rel_err_ll)

LL_RTOL = 1e-9  # BASELINE.json north_star: per-point log-likelihood within 1e-9 relative
P_RTOL = 1e-12  # SURVEY.md section 8(d) parity gates: per-bin p_j where p_j > 1e-300


def _model(case):
    return orc.Model(case['model'], case['k'], case['r'], case_hist(case), case['tail'],
                     **case_ctor_kwargs(case))


@pytest.mark.parametrize('name', golden_case_names())
def test_emulated_loglik_matches_reference(name):
    case = load_case(name)
    m = _model(case)
    got = emu.loglik_batch(m, case['points'])
    keep = ~emu.marked(got)  # marked points: the GPU's term-by-term kernel (tests/test_gpu_big_golden.py)
    assert keep.mean() >= 0.9
    rel = rel_err_ll(got[keep], np.array(case['ll'], dtype=float)[keep])
    assert rel.max() <= LL_RTOL, (name, int(rel.argmax()), np.array(case['points'])[keep][int(rel.argmax())])


@pytest.mark.parametrize('name', golden_case_names())
def test_emulated_probabilities_match_reference(name):
    case = load_case(name)
    m = _model(case)
    for idx, want in case['probs'].items():
        point = list(case['points'][int(idx)])
        _, got = emu.loglik_batch(m, [point], clip=True, want_probs=True)
        want = np.array(want, dtype=float)
        ok = want > 1e-300
        rel = np.abs(got[0][ok] - want[ok]) / want[ok]
        assert rel.max() <= P_RTOL, (name, idx, rel.max())
        # below the double range the reference returns exact zeros; so must we
        assert np.all(got[0][want == 0] == 0)


def test_single_term_sweep_against_oracle():
    """S = 1, a_0 = 1: the mixture is truncated_poisson itself.  Sweeps the rate over the
    reference's quirky regions: the `l > 1e-8` switch, the 2^-63 quantisation of expl(l) - 1 and
    the staged e^200 division."""
    hist = {j: 1 for j in range(1, 700)}
    m = orc.Model('basic', 21, 100, hist, 0, max_error=1)
    m._bounds[:] = np.nan
    rng = np.random.default_rng(11)
    rates = np.concatenate([
        np.exp(rng.uniform(np.log(1e-13), np.log(1e-3), 150)),
        np.exp(rng.uniform(np.log(1e-3), np.log(640), 150)),
        200.0 * rng.integers(1, 4, 60) + np.exp(rng.uniform(np.log(1e-9), np.log(50), 60)),
        [1e-8, 1.0000001e-8, 200.0, 200.001, 400.0, 400.00000001, 600.0000001, 2.0 ** -11],
    ])
    pts = [[float(L) / 0.8, 0.0] for L in rates]
    _, got = emu.loglik_batch(m, pts, clip=False, want_probs=True)
    for i in range(len(pts)):
        want = m.probs(pts[i])
        ok = want > 1e-300
        rel = np.abs(got[i][ok] - want[ok]) / want[ok]
        assert rel.max() <= P_RTOL, (rates[i], rel.max(), int(rel.argmax()))


def test_random_points_against_oracle_small_hist():
    case = load_case('cfg2_repeats')
    m = _model(case)
    rng = np.random.default_rng(3)
    n = 300
    pts = np.column_stack([30 * 3 ** rng.uniform(-1, 1, n), np.exp(rng.uniform(np.log(1e-4), np.log(.5), n)),
                           rng.uniform(.3, 1, n), rng.uniform(0, 1, n), rng.uniform(.02, 1, n)])
    got = emu.loglik_batch(m, pts)
    want = m.loglik_batch(pts, threads=8)
    keep = ~emu.marked(got)
    assert keep.mean() >= 0.97
    assert rel_err_ll(got[keep], want[keep]).max() <= LL_RTOL


def test_q1_one_is_the_basic_model():
    case = load_case('e05_basic')
    hist = case_hist(case)
    basic = orc.Model('basic', 21, 100, hist, 0, max_error=8)
    rep = orc.Model('repeats', 21, 100, hist, 0, max_error=8)
    a = emu.loglik_batch(basic, [[10, .05], [7, .01]])
    b = emu.loglik_batch(rep, [[10, .05, 1, .3, .4], [7, .01, 1, 0, 0]])
    assert rel_err_ll(a, b).max() <= 1e-14


def test_nan_and_empty_terms():
    case = load_case('e05_repeats')
    m = _model(case)
    got = emu.loglik_batch(m, [[math.nan, .05, .5, .5, .5]])
    assert math.isnan(got[0])
