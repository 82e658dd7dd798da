#!/usr/bin/env python
"""First-light measurement on a B200: FP64 peak probes and a quick timing of the likelihood
kernel on a cfg3-shaped batch (development aid; bench.py is the reported measurement)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from oracle import covest_oracle as orc  # noqa: E402
from tests.helpers import case_ctor_kwargs, case_hist, context_for, load_case  # noqa: E402

case = load_case('cfg3_repeats_dense1000')
m = orc.Model(case['model'], case['k'], case['r'], case_hist(case), case['tail'], **case_ctor_kwargs(case))
out = {}
with context_for(m) as ctx:
    out['sm'] = ctx.sm_count
    out['dfma_tflops'] = ctx.fp64_peak(0)
    out['dmma_tflops'] = ctx.fp64_peak(1)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    axes = [np.geomspace(10, 90, 40), np.geomspace(.003, .5, 25), np.linspace(.3, 1, 10),
            np.linspace(0, 1, 10), np.linspace(.05, 1, 10)]
    ctx.set_timing(True)
    total = 40 * 25 * 10 * 10 * 10
    stride = max(1, total // n)
    for rep in range(3):
        t = time.time()
        ll, rows = ctx.lattice_eval(axes, first=rep, stride=stride, count=n, k_best=8)
        wall = time.time() - t
        ms, launches = ctx.last_kernel_ms()
        out['run%d' % rep] = dict(points=n, kernel_ms=ms, wall_ms=wall * 1e3,
                                  point_bins_per_s=n * 1000 / (ms * 1e-3), finite=int(np.isfinite(ll).sum()))
    out['best_row'] = rows[0].tolist()
print(json.dumps(out, indent=1))
