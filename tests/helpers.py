"""Shared loaders for the golden fixtures (tests/golden/, written by gen_golden.py)."""
import glob
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_case_names():
    return sorted(os.path.basename(p)[len('loglik_'):-len('.json')]
                  for p in glob.glob(os.path.join(GOLDEN, 'loglik_*.json')))


def load_case(name):
    with open(os.path.join(GOLDEN, 'loglik_%s.json' % name)) as f:
        return json.load(f)


def load_tp_kat():
    with open(os.path.join(GOLDEN, 'tp_kat.json')) as f:
        return json.load(f)


def case_hist(case):
    return {int(j): int(h) for j, h in case['hist']}


def case_ctor_kwargs(case):
    kw = dict(max_error=case['max_error'], max_cov=case.get('max_cov'))
    if case['model'] == 'repeats':
        kw['min_single_copy_ratio'] = case.get('min_q1', 0.3)
        kw['threshold'] = case.get('threshold', 1e-8)
    return kw


def rel_err_ll(got, want):
    """Relative error per point; equal infinities / NaNs count as exact."""
    import numpy as np
    got = np.asarray(got, dtype=float)
    want = np.asarray(want, dtype=float)
    with np.errstate(all='ignore'):
        rel = np.abs(got - want) / np.abs(want)
    same = (np.isnan(got) & np.isnan(want)) | (np.isinf(got) & np.isinf(want) & (got == want))
    rel[same] = 0.0
    rel[np.isnan(rel)] = np.inf
    rel[(got == want)] = 0.0
    return rel


def context_for(oracle_model, device=None):
    """A device LikelihoodContext holding the same model description as an oracle Model."""
    from covest_b200.engine import LikelihoodContext
    m = oracle_model
    return LikelihoodContext(m.kind, m.k, m.r, m.max_error, m.bin_j, m.bin_h, m.tail, m.threshold,
                             m.bounds, m.comb, device=device)
