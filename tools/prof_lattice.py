#!/usr/bin/env python
"""The cfg3 / cfg5 benchmark lattice through cvb_lattice_eval a few times with phase timing: the
command the ncu captures of round 2 are taken on (development aid).

    python tools/prof_lattice.py [cfg3|cfg5] [repeats]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covest_b200 import workload  # noqa: E402
from covest_b200.models import RepeatsModel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = workload.CONFIGS[name]
hist = workload.synthetic_histogram(name)
model = RepeatsModel(cfg['k'], cfg['r'], hist, 0, max_error=8)
ctx = model.device_context
axes = workload.lattice_axes(cfg['theta'], n_c=40, n_e=25) if name == 'cfg3' else \
    workload.lattice_axes(cfg['theta'], n_c=25, n_e=50, n_q1=10, n_q2=10, n_q=100)
count = int(np.prod([len(a) for a in axes]))
out = torch.empty(count, dtype=torch.float64, device='cuda')
rows = torch.empty((64, 6), dtype=torch.float64, device='cuda')
ctx.set_timing(True)
for i in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ctx.lattice_eval(axes, out_ll=out, out_rows=rows, k_best=64)
    b.record()
    torch.cuda.synchronize()
    info = ctx.last_path_info()
    print('step %.3f ms, evaluation %.3f ms' % (a.elapsed_time(b), ctx.last_kernel_ms()[0]),
          {k: (round(v, 3) if isinstance(v, float) else v) for k, v in info.items()}, flush=True)
