#!/usr/bin/env python
"""Phase timing of the cfg3 lattice through the device-resident entry point (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from covest_b200 import workload  # noqa: E402
from covest_b200.models import RepeatsModel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
cfg = workload.CONFIGS[name]
hist = workload.synthetic_histogram(name)
model = RepeatsModel(cfg['k'], cfg['r'], hist, 0, max_error=8)
ctx = model.device_context
axes = workload.lattice_axes(cfg['theta'], n_c=40, n_e=25)
pts = torch.from_numpy(workload.lattice_points(axes)).cuda()
out = torch.empty(len(pts), dtype=torch.float64, device='cuda')
ctx.set_timing(True)
res = {}
for mode in (ctx.PATH_FACTORED_GEMM, ctx.PATH_FACTORED_PREFIX):
    ctx.set_path(mode)
    for i in range(4):
        ctx.loglik(pts, out=out)
        torch.cuda.synchronize()
        info = ctx.last_path_info()
        print('total %.3f ms' % ctx.last_kernel_ms()[0], {k: (round(v, 3) if isinstance(v, float) else v) for k, v in info.items()})
    res[mode] = out.clone()
a, b = res[ctx.PATH_FACTORED_GEMM], res[ctx.PATH_FACTORED_PREFIX]
fin = torch.isfinite(a)
print('kernels agree to', float(((a[fin] - b[fin]).abs() / a[fin].abs()).max()), 'non-finite equal', bool((torch.isfinite(b) == fin).all()))
t0 = torch.cuda.Event(enable_timing=True)
t1 = torch.cuda.Event(enable_timing=True)
t0.record()
rows = ctx.topk(out, pts, 64)
t1.record()
torch.cuda.synchronize()
print('topk %.3f ms' % t0.elapsed_time(t1), 'nonzero bins', int(sum(1 for v in hist.values() if v)), 'of', len(hist))
