#!/usr/bin/env python
"""Generate tests/golden/big_<cfg>.npz: the pinned C oracle (oracle/covest_oracle.c, bit-exact
against the unmodified reference on tests/golden/loglik_*.json, tests/test_oracle.py) evaluated on
the 10^4 seeded parity points of every BASELINE.json config (tests/bigpoints.py).

Offline job (minutes on 8 cores); the GPU tests only read the result.

    python tests/golden/gen_big_golden.py [cfg1 cfg2 ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import covest_oracle as orc  # noqa: E402
from tests import bigpoints  # noqa: E402


def histogram(cfg):
    """h_j ~ Poisson(N p_j(theta*)), j = 1..bins, zero counts kept; p_j from
    the oracle."""
    probe = orc.Model(cfg['model'], cfg['k'], cfg['r'], {j: 1 for j in range(1, cfg['bins'] + 1)}, 0, max_error=8)
    p = probe.probs(list(cfg['theta']))
    h = np.random.default_rng(cfg['seed']).poisson(cfg['kmers'] * np.maximum(p, 0.0))
    return np.arange(1, cfg['bins'] + 1, dtype=np.int32), h.astype(np.int64)


def main(names):
    threads = os.cpu_count() or 1
    for name in names:
        cfg = bigpoints.BIG_CONFIGS[name]
        j, h = histogram(cfg)
        tail = 0
        if cfg.get('trim'):  # histogram.py:111-134 trim_hist: bins below the cut stay, the rest is the tail
            keep = j < cfg['trim']
            tail = int(h[~keep].sum())
            j, h = j[keep], h[keep]
        pts = bigpoints.big_points(name)
        m = orc.Model(cfg['model'], cfg['k'], cfg['r'], dict(zip(j.tolist(), h.tolist())), tail, max_error=8)
        t0 = time.time()
        ll = m.loglik_batch(pts, mode=orc.LADDER, threads=threads)
        extra = {}
        if tail:
            # 1 - min(1, fsum(p_j)) of every point (models.py:103): log-likelihoods of a histogram with
            # a tail are ill-conditioned where the model puts (nearly) all mass inside the histogram
            import math
            omm = np.empty(len(pts))
            for i, p in enumerate(pts):
                q = list(p)
                for a, (lo, hi) in enumerate(m.bounds):
                    if lo is not None and q[a] < lo:
                        q[a] = lo
                    elif hi is not None and q[a] > hi:
                        q[a] = hi
                omm[i] = 1.0 - min(1.0, math.fsum(m.probs(q)))
            extra['one_minus_mass'] = omm
        np.savez_compressed(bigpoints.big_path(name), hist_j=j, hist_h=h, ll=ll, tail=np.array(tail),
                            points_sha256=np.array(bigpoints.points_digest(pts)), **extra)
        fin = np.isfinite(ll)
        print('%s: %d points x %d bins in %.0f s; finite %d, -inf %d, +inf/nan %d' % (
            name, len(pts), len(j), time.time() - t0, fin.sum(), np.isneginf(ll).sum(),
            (~fin & ~np.isneginf(ll)).sum()), flush=True)


if __name__ == '__main__':
    main(sys.argv[1:] or list(bigpoints.BIG_CONFIGS))
