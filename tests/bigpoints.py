"""The 10^4 seeded parity points per BASELINE.json config (parity gates) and
the loader of their oracle values (tests/golden/big_<cfg>.npz, written by
tests/golden/gen_big_golden.py with the pinned C oracle).

The points are regenerated from the seed (numpy's default_rng stream is stable) and checked against
the digest stored next to the values, so the fixtures hold only the histogram and the 10^4
log-likelihoods.

Per config: uniform-random points in the box of initial_grid (covest/grid.py:95-98) including the
small-q region q in [0.02, 0.3] where the copy series is long, points of the benchmark lattice
itself (cfg3, cfg5), rates just above 200 n (the staged division of
c_src/covest_poissonmodule.c:25-28), and the bound / edge points err = 0, q1 = 1, q2 = 0, q = 0,
q = 1 and out-of-bounds arguments (models.py:60-69).
"""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
N_BIG = 10000

# (model, k, r, theta*, bins, distinct k-mers, seed): covest_b200/workload.py CONFIGS, restated here
# so that the fixtures do not depend on the package
BIG_CONFIGS = {
    'cfg1': dict(model='basic', k=21, r=100, theta=(10.0, 0.03), bins=300, kmers=1e7, seed=1001),
    'cfg2': dict(model='repeats', k=21, r=100, theta=(30.0, 0.03, 0.7, 0.5, 0.5), bins=300, kmers=1e7, seed=1002),
    'cfg3': dict(model='repeats', k=21, r=100, theta=(30.0, 0.03, 0.7, 0.5, 0.5), bins=1000, kmers=1e8, seed=1003),
    'cfg4': dict(model='repeats', k=31, r=150, theta=(200.0, 0.01, 0.7, 0.5, 0.28), bins=5000, kmers=1e7, seed=1004),
    'cfg5': dict(model='repeats', k=21, r=100, theta=(30.0, 0.03, 0.7, 0.5, 0.5), bins=2000, kmers=1e8, seed=1005),
    # cfg2 trimmed at 30 with the rest of the counts as the tail (models.py:103-104: the mass term)
    'cfg2t': dict(model='repeats', k=21, r=100, theta=(30.0, 0.03, 0.7, 0.5, 0.5), bins=300, kmers=1e7, seed=1002,
                  trim=30),
}


def lattice_axes(name):
    """The benchmark lattices of bench.py (cfg3: 40 x 25 x 10 x 10 x 10 on one rank; cfg5:
    200 x 50 x 10 x 10 x 100 = 10^8 points over 8 ranks)."""
    c, e = BIG_CONFIGS[name]['theta'][:2]
    n = {'cfg3': (40, 25, 10, 10, 10), 'cfg5': (200, 50, 10, 10, 100)}[name]
    return [np.geomspace(c / 3, 3 * c, n[0]), np.geomspace(e / 3, min(0.5, 3 * e), n[1]),
            np.linspace(0.3, 1.0, n[2]), np.linspace(0.0, 1.0, n[3]), np.linspace(0.05, 1.0, n[4])]


def _lattice_rows(axes, idx):
    cols = []
    idx = np.asarray(idx, dtype=np.int64)
    for a in reversed(axes):
        cols.append(np.asarray(a, dtype=np.float64)[idx % len(a)])
        idx = idx // len(a)
    return np.column_stack(cols[::-1])


def big_points(name, n=N_BIG):
    cfg = BIG_CONFIGS[name]
    rng = np.random.default_rng(900000 + cfg['seed'])
    c0, k, r = cfg['theta'][0], cfg['k'], cfg['r']
    hi = 0.3 if name == 'cfg4' else 1.0  # cfg4: 3 * 200 * 49 copies stays below the reference's overflow (~11 360)
    cols = [c0 * 3 ** rng.uniform(-1, hi, n), np.exp(rng.uniform(np.log(1e-4), np.log(.5), n))]
    if cfg['model'] == 'basic':
        pts = np.column_stack(cols)
        pts[0] = [c0, 0.0]
        pts[1] = [0.001, 0.03]   # below the coverage bound
        pts[2] = [c0, 0.9]       # above the error bound
        pts[3] = [c0, 0.5]
        pts[4] = [c0, 1e-12]
        # rates just above 200 n (only reachable by coverage itself in the basic model)
        for i, (nn, d) in enumerate([(1, 1e-9), (1, 1e-3), (1, .5), (2, 1e-6), (2, 5.0), (1, -1e-6)]):
            e = 0.01
            pts[5 + i] = [(200 * nn + d) / (1 - e) ** k * r / (r - k + 1), e]
        return np.ascontiguousarray(pts)
    q_lo = 0.05 if name == 'cfg4' else 0.02
    if name == 'cfg2t':  # stay where the histogram carries most of the mass: the tail term is the point here
        cols = [c0 * 3 ** rng.uniform(-0.5, 0.5, n), np.exp(rng.uniform(np.log(3e-3), np.log(.2), n))]
    q = np.where(rng.uniform(size=n) < 0.4, rng.uniform(q_lo, 0.3, n), rng.uniform(0.3, 1, n))
    cols += [rng.uniform(.3, 1, n), rng.uniform(0, 1, n), q]
    pts = np.column_stack(cols)
    at = 0
    if name in ('cfg3', 'cfg5'):  # points of the benchmark lattice itself, every q value included
        axes = lattice_axes(name)
        total = int(np.prod([len(a) for a in axes]))
        m = n // 2
        pts[:m] = _lattice_rows(axes, rng.choice(total, m, replace=False))
        at = m
    # rates o * l_0 just above (and once just below) 200 n
    saw = [(o, nn, d) for o in (4, 7, 12, 25) for nn in (1, 2) for d in (1e-9, 1e-6, 1e-3, .5, 5.0, -1e-6)]
    if name == 'cfg4':
        saw = [(o, nn, d) for o in (1, 2, 3, 10) for nn in (1, 3, 9) for d in (1e-9, 1e-6, 1e-3, .5, 5.0, -1e-6)]
    for i, (o, nn, d) in enumerate(saw):
        e = float(np.exp(rng.uniform(np.log(1e-3), np.log(.05))))
        c = (200 * nn + d) / o / (1 - e) ** k * r / (r - k + 1)
        pts[at + i, :2] = [c, e]
        pts[at + i, 4] = rng.uniform(0.1, 0.6)
    at += len(saw)
    edges = [[c0, 0.0, .8, .5, .5], [c0, .03, 1.0, .5, .5], [c0, .03, .7, 0.0, .5], [c0, .03, .7, .5, 0.0],
             [c0, .03, .7, .5, 1.0], [c0, .9, .1, 2, -1], [c0, .03, .3, 1.0, .5], [c0, .03, 1.0, 0.0, 0.0],
             [0.001, .03, .7, .5, .5], [c0, .5, .3, 0.0, 1.0], [c0, 1e-12, .7, .5, .5], [c0 / 3, .03, 1.0, 0, .05],
             [3 * c0 if name != 'cfg4' else c0, .1, 1.0, 0, .05], [c0 / 3, .1, .3, 0.0, .05]]
    for i, row in enumerate(edges):
        pts[at + i] = row
    return np.ascontiguousarray(pts)


def points_digest(pts):
    return hashlib.sha256(np.ascontiguousarray(pts, dtype=np.float64).tobytes()).hexdigest()


def big_path(name):
    return os.path.join(GOLDEN, 'big_%s.npz' % name)


def load_big(name):
    """-> dict(cfg, hist {j: h}, tail, points (10^4 x n_param), ll (oracle), digest-checked)."""
    with np.load(big_path(name)) as z:
        hist = {int(j): int(h) for j, h in zip(z['hist_j'], z['hist_h'])}
        ll = z['ll'].astype(np.float64)
        digest = str(z['points_sha256'])
        tail = int(z['tail']) if 'tail' in z else 0
        omm = z['one_minus_mass'].astype(np.float64) if 'one_minus_mass' in z else None
    pts = big_points(name, len(ll))
    if points_digest(pts) != digest:
        raise RuntimeError('big_points(%r) no longer reproduces the points the fixture was made on' % name)
    return dict(cfg=BIG_CONFIGS[name], hist=hist, tail=tail, points=pts, ll=ll, one_minus_mass=omm)


MASS_ULPS = 8


def ll_tolerance(big, rtol=1e-9):
    """Per-point absolute tolerance of the log-likelihood: rtol * |ll|, plus -- for a histogram with a
    tail -- what MASS_ULPS units in the last place of sp = fsum(p_j) move the tail term
    tail * log(1 - sp) (models.py:103-104).  Where the model puts all but 1e-9 of the mass inside the
    histogram, ONE ulp of sp (1.1e-16, i.e. one ulp of any leading p_j) moves the reference's own value
    by more than 1e-9 relative (measured: 2.6e-8 at 1 - sp = 2.3e-11): no implementation that is not
    bit-identical in every rounding, libm's pow included, can meet the plain gate there.  The extra
    term is below 1e-10 * |ll| wherever 1 - sp > 1e-7."""
    tol = rtol * np.abs(big['ll'])
    if big['tail'] and big['one_minus_mass'] is not None:
        omm = big['one_minus_mass']
        with np.errstate(divide='ignore'):  # sp >= 1: no tail term in the reference (models.py:104), plain gate
            tol = tol + np.where(omm > 0, big['tail'] * MASS_ULPS * 2.0 ** -53 / omm, 0.0)
    return tol
