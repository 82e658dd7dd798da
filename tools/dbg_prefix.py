import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from covest_b200 import workload
from covest_b200.models import RepeatsModel
cfg = workload.CONFIGS['cfg3']
hist = workload.synthetic_histogram('cfg3')
model = RepeatsModel(cfg['k'], cfg['r'], hist, 0, max_error=8)
ctx = model.device_context
axes = workload.lattice_axes(cfg['theta'], n_c=40, n_e=25)
hp = workload.lattice_points(axes)
pts = torch.from_numpy(hp).cuda()
out = torch.empty(len(pts), dtype=torch.float64, device='cuda')
res = {}
for mode in (ctx.PATH_FACTORED_GEMM, ctx.PATH_FACTORED_PREFIX):
    ctx.set_path(mode)
    ctx.loglik(pts, out=out); torch.cuda.synchronize()
    res[mode] = out.cpu().numpy().copy()
a, b = res[3], res[4]
bad = np.nonzero(~((a == b) | (np.abs(a - b) <= 1e-11 * np.abs(a))))[0]
print('bad', len(bad))
for i in bad[:20]:
    print(i, hp[i], a[i], b[i])
# the group of the first bad point alone, through each path
g = bad[0] // 1000
sub = torch.from_numpy(hp[g * 1000:(g + 1) * 1000]).cuda()
k = bad[0] - g * 1000
for mode, name in ((1, 'per-point'), (3, 'gemm'), (4, 'prefix')):
    ctx.set_path(mode)
    r = ctx.loglik(sub).cpu().numpy() if hasattr(ctx.loglik(sub), 'cpu') else np.asarray(ctx.loglik(sub))
    print(name, 'group alone', repr(float(r[k])), ctx.last_path_info()['kernel'])
ctx.set_path(3)
for n in (128, 256, 1000):
    for rep in range(2):
        r = np.asarray(ctx.loglik(hp[g * 1000 + 500:g * 1000 + 500 + n]))
        print('gemm', n, 'points from 500, rep', rep, repr(float(r[k - 500])))
a2 = np.asarray(ctx.loglik(hp))
print('gemm full again: same bad set', np.array_equal(np.nonzero(a2 != b)[0], np.nonzero(a != b)[0]), int((a2 != a).sum()))
from oracle import covest_oracle as orc
om = orc.Model('repeats', cfg['k'], cfg['r'], {int(j): int(v) for j, v in hist.items()}, 0, max_error=8)
sel = bad[::10][:6]
ref = om.loglik_batch(hp[sel], threads=8)
for i, r in zip(sel, ref):
    print('oracle on the device histogram', i, repr(float(r)), 'gemm', repr(float(a[i])), 'prefix', repr(float(b[i])))
