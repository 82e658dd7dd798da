#!/usr/bin/env python
"""Parity statistics of the device paths on the 10^4-point oracle goldens (tests/bigpoints.py):
per config and path the number of points beyond 1e-9, infinity mismatches and the worst point.
Development aid; the gate itself is tests/test_gpu_big_golden.py.

    python tools/parity_report.py [cfg ...] > gpurun_out/parity.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import bigpoints  # noqa: E402
from tests.helpers import rel_err_ll  # noqa: E402
from tests.test_gpu_big_golden import _model  # noqa: E402

out = []
for name in sys.argv[1:] or ['cfg1', 'cfg2', 'cfg2t', 'cfg3', 'cfg4', 'cfg5']:
    big = bigpoints.load_big(name)
    want = big['ll']
    inside = ~(np.isposinf(want) | np.isnan(want))
    model = _model(big)
    ctx = model.device_context
    for path in ([0, 5] if big['cfg']['model'] == 'basic' else [0, 1, 3, 4, 5]):
        ctx.set_path(path)
        t0 = time.time()
        got = ctx.loglik(big['points'])
        dt = time.time() - t0
        rel = rel_err_ll(got[inside], want[inside])
        with np.errstate(invalid='ignore'):
            err = np.where(rel == 0, 0.0, np.abs(got[inside] - want[inside]))
        bad = np.nonzero(~(err <= bigpoints.ll_tolerance(big)[inside]))[0]  # 1e-9 |ll| (+ mass ulps with a tail)
        fin = np.isfinite(rel)
        rec = {'cfg': name, 'path': path, 'kernel': ctx.last_path_info()['kernel'], 'seconds': dt,
               'refined': ctx.last_path_info()['refined_points'],
               'inside': int(inside.sum()), 'bad': int(len(bad)), 'beyond_1e-9': int((~(rel <= 1e-9)).sum()), 'inf_mismatch': int((~fin).sum()),
               'max_finite_rel': float(rel[fin].max()) if fin.any() else None,
               'outside_finite_on_device': int(np.isfinite(got[~inside]).sum()),
               'examples': [{'point': big['points'][inside][i].tolist(), 'got': float(got[inside][i]),
                             'want': float(want[inside][i]), 'index': int(np.nonzero(inside)[0][i])}
                            for i in bad[:12]]}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    model.close()

if '--lattice' in os.environ.get('PARITY_EXTRA', ''):
    # the cfg3 benchmark lattice: where do the device paths disagree, and what does the oracle say there?
    from covest_b200 import workload
    from covest_b200.models import RepeatsModel
    from oracle import covest_oracle as orc
    big = bigpoints.load_big('cfg3')
    cfg = big['cfg']
    model = RepeatsModel(cfg['k'], cfg['r'], big['hist'], 0, max_error=8)
    ctx = model.device_context
    pts = workload.lattice_points(bigpoints.lattice_axes('cfg3'))
    vals = {}
    for path in (0, 1, 3):
        ctx.set_path(path)
        vals[path] = ctx.loglik(pts).copy()
    model.close()
    rel = rel_err_ll(vals[0], vals[1])
    off = np.nonzero(~(rel <= 1e-11))[0]
    m = orc.Model('repeats', cfg['k'], cfg['r'], big['hist'], 0, max_error=8)
    pick = off[:400]
    want = m.loglik_batch(pts[pick], threads=os.cpu_count() or 1)
    rows = []
    for i, w in zip(pick, want):
        rows.append({'index': int(i), 'point': pts[i].tolist(), 'oracle': float(w),
                     'prefix': float(vals[0][i]), 'direct': float(vals[1][i]), 'gemm': float(vals[3][i])})
    rec = {'lattice': 'cfg3', 'disagree_prefix_vs_direct': int(len(off)),
           'neginf_prefix': int(np.isneginf(vals[0]).sum()), 'neginf_direct': int(np.isneginf(vals[1]).sum()),
           'rows': rows}
    print(json.dumps(rec), flush=True)
