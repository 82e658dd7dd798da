"""End to end on the device: the reference's estimator flow (process_histogram -> model ->
CoverageEstimator.compute_coverage, covest/covest.py:41-96) on the synthetic cfg1 / cfg2 histograms
of BASELINE.json, against results produced by the unmodified reference
(tests/golden/gen_e2e_golden.py).

Stock L-BFGS-B differentiates with 1e-8 forward steps and stops ~1e-4 (relative) short of the
optimum, at a point that depends on rounding noise; parity of the *estimates* is therefore
checked at the optimum both implementations approach: the Newton
polish of each one's own objective.  Tolerance: 1e-6 relative on coverage and genome size
(BASELINE.json north_star).  Needs a B200."""
import io
import json
import os
import time
from contextlib import redirect_stdout

import numpy as np
import pytest
import yaml

from covest_b200.covest import CoverageEstimator
from covest_b200.histogram import process_histogram
from covest_b200.models import select_model
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu

EST_RTOL = 1e-6  # BASELINE.json north_star: final coverage and genome_size


def _golden():
    with open(os.path.join(GOLDEN, 'e2e_golden.json')) as f:
        return json.load(f)


def _genome_size(model, hist, coverage):
    kmers = sum(j * h for j, h in hist.items())
    return kmers / model.correct_c(coverage)


@pytest.mark.parametrize('name', ['cfg1_basic', 'cfg2_repeats'])
def test_estimates_match_the_reference_at_the_polished_optimum(name):
    g = _golden()[name]
    hist = {int(j): int(h) for j, h in g['hist']}
    h2, tail, sf, gc, ge = process_histogram(hist, g['k'], g['r'], **g['flags'])
    assert [[j, h] for j, h in h2.items()] == g['processed']['hist'] and tail == g['processed']['tail']
    model = select_model(g['model'])(g['k'], g['r'], h2, tail, max_error=8)
    guess = list(model.defaults)
    guess[:2] = gc, ge
    assert guess == pytest.approx(g['guess'], rel=1e-13)
    est = CoverageEstimator(model)
    t0 = time.perf_counter()
    x, ok = est.compute_coverage(guess)
    wall = time.perf_counter() - t0
    assert ok
    # the stock runs end within the optimiser's own noise floor of each other ...
    assert x[0] == pytest.approx(g['stock']['x'][0], rel=2e-3)
    # (one ulp of the objective moves the end point of L-BFGS-B's 1e-8
    # forward differences by 6e-4 in coverage; along the flat valley that is ~1e-7 of the value)
    assert -est.likelihood_f(x) == pytest.approx(g['stock']['loglikelihood'], rel=1e-6)
    # ... and at the same optimum once polished
    xp, fp = est.polish(x)
    want = g['polished']['x']
    assert xp[0] == pytest.approx(want[0], rel=EST_RTOL)
    assert _genome_size(model, hist, xp[0]) == pytest.approx(_genome_size(model, hist, want[0]), rel=EST_RTOL)
    assert xp[1] == pytest.approx(want[1], rel=1e-5)
    assert fp == pytest.approx(g['polished']['objective'], rel=1e-11)
    print('%s: device estimate in %.3f s (reference on the build container: %.3f s), %d launches' % (
        name, wall, g['stock']['wall_s'], est.launches))


def test_cli_report_has_the_reference_keys(tmp_path):
    """`covest hist -m repeat -k 21 -r 100 -sf 1 --polish`: YAML keys of covest/data.py:134-171."""
    from covest_b200 import covest as cli
    g = _golden()['cfg2_repeats']
    path = tmp_path / 'cfg2.hist'
    path.write_text(''.join('%d %d\n' % (j, h) for j, h in g['hist']))
    buf = io.StringIO()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with redirect_stdout(buf):
            cli.run([str(path), '-m', 'repeat', '-k', '21', '-r', '100', '-sf', '1', '--polish', '--seed', '1'])
    finally:
        os.chdir(cwd)
    out = yaml.safe_load(buf.getvalue())
    for key in ('model', 'hist_size', 'sample_factor', 'success', 'coverage', 'error_rate', 'q1', 'q2', 'q',
                'loglikelihood', 'genome_size', 'guessed_coverage', 'guessed_error_rate'):
        assert key in out, (key, sorted(out))
    assert out['coverage'] == pytest.approx(g['polished']['x'][0], rel=EST_RTOL)
    hist = {int(j): int(h) for j, h in g['hist']}
    kmers = sum(j * h for j, h in hist.items())
    want_size = kmers / (g['polished']['x'][0] * (100 - 21 + 1) / 100)
    assert out['genome_size'] == pytest.approx(want_size, rel=EST_RTOL)


def test_cli_lattice_starts(tmp_path):
    """`--lattice-starts K`: the K best points of a candidate lattice as starts, lock-step refinement."""
    from covest_b200 import covest as cli
    g = _golden()['cfg2_repeats']
    path = tmp_path / 'cfg2.hist'
    path.write_text(''.join('%d %d\n' % (j, h) for j, h in g['hist']))
    buf = io.StringIO()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with redirect_stdout(buf):
            cli.run([str(path), '-m', 'repeat', '-k', '21', '-r', '100', '-sf', '1', '--lattice-starts', '16'])
    finally:
        os.chdir(cwd)
    out = yaml.safe_load(buf.getvalue())
    assert out['success']
    # at least as good as the optimum the polished reference run ends at
    assert out['loglikelihood'] >= -g['polished']['objective'] * (1 + 1e-12)
    if out['loglikelihood'] <= -g['polished']['objective'] * (1 - 1e-10):
        assert out['coverage'] == pytest.approx(g['polished']['x'][0], rel=EST_RTOL)
