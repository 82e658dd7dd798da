#!/usr/bin/env python
"""End-to-end golden results produced by the UNMODIFIED reference (build container only; needs
/root/reference): the stock estimator flow on the synthetic cfg1 / cfg2 histograms of
BASELINE.json, its wall time on this container's CPU, and the *polished* optimum of the
reference's own objective (stock L-BFGS-B stops ~1e-4 away from the
optimum because of its 1e-8 forward differences, so end-to-end parity is defined at the optimum
both implementations approach).  Writes tests/golden/e2e_golden.json.

    python tests/golden/gen_e2e_golden.py
"""
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden  # noqa: E402  (reuses the import shims)


def polish(f, x, lo, hi, iterations=30, rel_step=1e-4, tol=1e-10):
    """Projected Newton on central differences of the reference objective f (sequential calls)."""
    x = np.array(x, dtype=float)
    n = len(x)
    for _ in range(iterations):
        h = rel_step * np.maximum(np.abs(x), 1e-3)
        free = [i for i in range(n) if x[i] - h[i] >= lo[i] and x[i] + h[i] <= hi[i]]
        if not free:
            break
        fx = f(x)
        g = np.zeros(len(free))
        H = np.zeros((len(free), len(free)))
        for a, i in enumerate(free):
            e = np.zeros(n)
            e[i] = h[i]
            fp, fm = f(x + e), f(x - e)
            g[a] = (fp - fm) / (2 * h[i])
            H[a, a] = (fp - 2 * fx + fm) / h[i] ** 2
        for a, i in enumerate(free):
            for b in range(a + 1, len(free)):
                j = free[b]
                ei = np.zeros(n)
                ej = np.zeros(n)
                ei[i] = h[i]
                ej[j] = h[j]
                v = (f(x + ei + ej) - f(x + ei - ej) - f(x - ei + ej) + f(x - ei - ej)) / (4 * h[i] * h[j])
                H[a, b] = H[b, a] = v
        step = np.linalg.solve(H, -g)
        if g @ step > 0:
            step = -g * h[free] ** 2
        new = x.copy()
        new[free] = np.clip(x[free] + step, lo[free], hi[free])
        moved = np.max(np.abs(new - x) / np.maximum(np.abs(x), 1e-300))
        x = new
        if moved < tol:
            break
    return x, f(x)


def run_case(name, model_name, flags):
    from covest import covest as ref_covest
    from covest import histogram
    from covest.models import select_model
    with open(os.path.join(HERE, 'loglik_%s.json' % name)) as f:
        case = json.load(f)
    hist = {int(j): int(h) for j, h in case['hist']}
    random.seed(2024)
    np.random.seed(2024)
    t0 = time.perf_counter()
    h2, tail, sf, gc, ge = histogram.process_histogram(hist, case['k'], case['r'], **flags)
    model = select_model(model_name)(case['k'], case['r'], h2, tail, max_error=8)
    guess = list(model.defaults)
    guess[:2] = gc, ge
    est = ref_covest.CoverageEstimator(model)
    res, ok = est.compute_coverage(guess)
    wall = time.perf_counter() - t0
    ll = model.compute_loglikelihood(*res)
    lo = np.array([-np.inf if b[0] is None else b[0] for b in model.bounds])
    hi = np.array([np.inf if b[1] is None else b[1] for b in model.bounds])
    xp, fp = polish(lambda x: -model.compute_loglikelihood(*x), res, lo, hi)
    genome = lambda c: float(sum(j * h for j, h in hist.items()) / (c * (case['r'] - case['k'] + 1) / case['r']))
    return dict(model=model_name, k=case['k'], r=case['r'], flags=flags, hist=[[j, h] for j, h in hist.items()],
                processed=dict(hist=[[j, h] for j, h in h2.items()], tail=tail, sample_factor=sf),
                guess=guess, stock=dict(x=[float(v) for v in res], success=bool(ok), loglikelihood=float(ll),
                                        wall_s=wall, cpu='build container, 1 process'),
                polished=dict(x=[float(v) for v in xp], objective=float(fp)))


def main():
    gen_golden._init()
    out = {'cfg1_basic': run_case('cfg1_basic', 'basic', dict(sample_factor=1)),
           'cfg2_repeats': run_case('cfg2_repeats', 'repeat', dict(sample_factor=1))}
    with open(os.path.join(HERE, 'e2e_golden.json'), 'w') as f:
        json.dump(out, f)
    for k, v in out.items():
        print(k, v['stock'], v['polished'])


if __name__ == '__main__':
    main()
