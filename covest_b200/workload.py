"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d): model-sampled histograms,
candidate lattices, and the algorithmic work they contain."""
import math

import numpy as np

from .models import BasicModel, RepeatsModel

# (model, k, r, theta*, bins, distinct k-mers, seed) -- SURVEY.md section 8(d)
CONFIGS = {
    'cfg1': dict(model='basic', k=21, r=100, theta=(10.0, 0.03), bins=300, kmers=1e7, seed=1001),
    'cfg2': dict(model='repeats', k=21, r=100, theta=(30.0, 0.03, 0.7, 0.5, 0.5), bins=300,
                 kmers=1e7, seed=1002),
    'cfg3': dict(model='repeats', k=21, r=100, theta=(30.0, 0.03, 0.7, 0.5, 0.5), bins=1000,
                 kmers=1e8, seed=1003),
    'cfg4': dict(model='repeats', k=31, r=150, theta=(200.0, 0.01, 0.7, 0.5, 0.28), bins=5000,
                 kmers=1e7, seed=1004),
    'cfg5': dict(model='repeats', k=21, r=100, theta=(30.0, 0.03, 0.7, 0.5, 0.5), bins=2000,
                 kmers=1e8, seed=1005),
}


def model_class(name):
    return RepeatsModel if name.startswith('r') else BasicModel


def synthetic_histogram(cfg, keep_zeros=True, max_error=8):
    """h_j ~ Poisson(N * p_j(theta*)) for j = 1..bins, p_j from the model itself evaluated on the
    device ("simulated from the model")."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    probe = model_class(c['model'])(c['k'], c['r'], {j: 1 for j in range(1, c['bins'] + 1)}, 0,
                                    max_error=max_error)
    p = probe.probabilities_batch(np.array([c['theta']], dtype=np.float64))[0]
    probe.close()
    rng = np.random.default_rng(c['seed'])
    h = rng.poisson(c['kmers'] * np.maximum(p, 0.0))
    return {j: int(v) for j, v in zip(range(1, c['bins'] + 1), h) if keep_zeros or v > 0}


def lattice_axes(theta, n_c=40, n_e=25, n_q1=10, n_q2=10, n_q=10):
    """The candidate box of initial_grid (covest/grid.py:95-98): coverage and error rate
    log-spaced over theta/3 .. 3*theta (error capped at its bound 0.5), q1 in 0.3..1, q2 in 0..1,
    q in 0.05..1."""
    c, e = theta[0], theta[1]
    return [np.geomspace(c / 3, 3 * c, n_c), np.geomspace(e / 3, min(0.5, 3 * e), n_e),
            np.linspace(0.3, 1.0, n_q1), np.linspace(0.0, 1.0, n_q2), np.linspace(0.05, 1.0, n_q)]


def lattice_points(axes, first=0, stride=1, count=None, block=1):
    """Explicit rows of a lattice slice, last axis fastest (itertools.product order): runs of
    `block` consecutive indices, run j starting at (first + j*stride)*block (the slicing of
    cvb_lattice_eval)."""
    lens = [len(a) for a in axes]
    total = int(np.prod(lens))
    if count is None:
        runs = (total + block - 1) // block
        mine = max(0, (runs - first + stride - 1) // stride)
        count = mine * block
        if mine and (first + (mine - 1) * stride + 1) * block > total:
            count -= (first + (mine - 1) * stride + 1) * block - total
    i = np.arange(count, dtype=np.int64)
    idx = (first + (i // block) * stride) * block + i % block
    cols = []
    for a, n in zip(reversed(axes), reversed(lens)):
        cols.append(np.asarray(a, dtype=np.float64)[idx % n])
        idx = idx // n
    return np.ascontiguousarray(np.column_stack(cols[::-1]))


def copy_cutoff(points, max_bin, threshold=1e-8):
    """O_thr of every row (models.py:185-191), vectorised."""
    pts = np.asarray(points, dtype=np.float64)
    q1, q2, q = pts[:, 2], pts[:, 3], pts[:, 4]
    two = (1 - q1) * q2
    many = (1 - q1) * (1 - q2) * q
    out = np.full(len(pts), max_bin, dtype=np.int64)
    if threshold is None:
        return out
    todo = np.ones(len(pts), dtype=bool)
    for o in range(1, max_bin):
        if o == 1:
            b = q1
        elif o == 2:
            b = two
        else:
            with np.errstate(all='ignore'):
                b = many * (1 - q) ** (o - 3)
        hit = todo & (b <= threshold)
        out[hit] = o
        todo &= ~hit
        if not todo.any():
            break
    return out


def term_counts(model, points):
    """Mixture terms T = S * (O_thr - 1) per point (S for the basic model)."""
    if not model.repeats:
        return np.full(len(points), model.max_error, dtype=np.int64)
    clipped = np.array([model.fit_to_bounds(list(p)) for p in np.asarray(points)], dtype=np.float64) \
        if len(points) < 4096 else _clip(model, np.asarray(points, dtype=np.float64))
    return model.max_error * np.maximum(copy_cutoff(clipped, max(model.hist), model.threshold) - 1, 0)


def _clip(model, pts):
    out = pts.copy()
    for i, (lo, hi) in enumerate(model.bounds):
        if lo is not None:
            out[:, i] = np.maximum(out[:, i], lo)
        if hi is not None:
            out[:, i] = np.minimum(out[:, i], hi)
    return out


# FP64 work of one evaluation, DESIGN.md section 6: one FMA (2 flop) per (mixture term, bin) plus
# 64 flop per bin for the epilogue (a log, the count-weighted sum, the mass).
FLOP_PER_TERM_BIN = 2.0
FLOP_PER_BIN = 64.0
PREFIX_FLOP_PER_BIN = 6.0  # per slot of a profile row: q1*P1 (1) + two*P2 (2) + many*R (2) + the add into the mass (1)


def algorithmic_flop(n_bins, terms):
    terms = np.asarray(terms, dtype=np.float64)
    return float(np.sum(n_bins * (FLOP_PER_TERM_BIN * terms + FLOP_PER_BIN)))


def factored_flop(model, points, n_bins, counted_bins, row_bins=None):
    """FP64 work of the factored path on a batch (row_bins: the slots of a
    profile row the kernels after K1 carry, cvb_last_path_info's row_slots; default every bin):
      profile_flop  one FMA per (error class, copy number, bin) for every copy number up to the
                    largest cut-off of each distinct (coverage, error_rate)
      gemm_flop     one FMA per (copy number below the point's cut-off, bin) for every point, plus
                    64 flop per (point, bin with a non-zero count): the multiply by the bin's
                    scale, the logarithm and the count-weighted accumulation of models.py:100-107
    """
    pts = _clip(model, np.asarray(points, dtype=np.float64))
    copies = np.maximum(copy_cutoff(pts, max(model.hist), model.threshold) - 1, 0).astype(np.float64)
    keys = np.ascontiguousarray(pts[:, :2]).view([('c', 'f8'), ('e', 'f8')]).ravel()
    _, inverse = np.unique(keys, return_inverse=True)
    gmax = np.zeros(inverse.max() + 1)
    np.maximum.at(gmax, inverse, copies)
    rb = n_bins if row_bins is None else row_bins
    profile = FLOP_PER_TERM_BIN * model.max_error * n_bins * float(gmax.sum())
    gemm = float(np.sum(FLOP_PER_TERM_BIN * rb * copies + FLOP_PER_BIN * counted_bins))
    # prefix kernel: per (point, bin) the three-term combination (a multiply and two FMAs) and the
    # add into the mass; per (point, bin with a count) the same 64 flop as above; per (q-run, copy
    # number beyond 2 up to the run's largest cut-off, bin) one FMA into the running sum
    rkeys = np.ascontiguousarray(pts[:, [0, 1, 4]]).view([('c', 'f8'), ('e', 'f8'), ('q', 'f8')]).ravel()
    _, rinv = np.unique(rkeys, return_inverse=True)
    rmax = np.zeros(rinv.max() + 1)
    np.maximum.at(rmax, rinv, copies)
    prefix = float(len(pts)) * (PREFIX_FLOP_PER_BIN * rb + FLOP_PER_BIN * counted_bins) + \
        FLOP_PER_TERM_BIN * rb * float(np.maximum(rmax - 2, 0).sum())
    return {'profile_flop': profile, 'gemm_flop': gemm, 'prefix_flop': prefix, 'groups': int(len(gmax)),
            'q_runs': int(len(rmax)), 'mean_copies': float(copies.mean())}


def lattice_term_stats(model, axes):
    """Sum and mean of T over a whole lattice without materialising it: T only depends on the
    (q1, q2, q) axes."""
    q_pts = lattice_points([np.array([1.0]), np.array([0.1])] + list(axes[2:]))
    t = term_counts(model, q_pts)
    reps = len(axes[0]) * len(axes[1])
    return float(t.sum()) * reps, float(t.mean())


def lattice_flop(model, axes, n_groups, n_bins, counted_bins, row_bins=None):
    """factored_flop of `n_groups` whole (coverage, error_rate) groups of the lattice of `axes`
    without materialising the points: cut-offs and q-runs depend on the (q1, q2, q) axes only, every
    group holds the same combinations."""
    rb = n_bins if row_bins is None else row_bins
    q_pts = lattice_points([np.array([1.0]), np.array([0.1])] + [np.asarray(a, dtype=np.float64) for a in axes[2:]])
    q_pts = _clip(model, q_pts)
    copies = np.maximum(copy_cutoff(q_pts, max(model.hist), model.threshold) - 1, 0).astype(np.float64)
    _, rinv = np.unique(q_pts[:, 4], return_inverse=True)
    rmax = np.zeros(rinv.max() + 1)
    np.maximum.at(rmax, rinv, copies)
    m = len(q_pts)
    profile = FLOP_PER_TERM_BIN * model.max_error * n_bins * float(copies.max()) * n_groups
    gemm = float(np.sum(FLOP_PER_TERM_BIN * rb * copies + FLOP_PER_BIN * counted_bins)) * n_groups
    prefix = n_groups * (m * (PREFIX_FLOP_PER_BIN * rb + FLOP_PER_BIN * counted_bins) +
                         FLOP_PER_TERM_BIN * rb * float(np.maximum(rmax - 2, 0).sum()))
    return {'profile_flop': profile, 'gemm_flop': gemm, 'prefix_flop': float(prefix), 'groups': int(n_groups),
            'q_runs': int(n_groups * len(rmax)), 'mean_copies': float(copies.mean()),
            'mean_terms': float(model.max_error * copies.mean()), 'sum_terms_per_group': float(model.max_error * copies.sum())}


def random_box_points(theta, n, seed):
    """Seeded uniform-random points of the box initial_grid draws from (covest/grid.py:95-110: every
    coordinate independently uniform in [v / 3, 3 v] cut to the bounds), with the q axes over their
    whole ranges as in lattice_axes.  No two points share (coverage, error_rate)."""
    rng = np.random.default_rng(seed)
    c, e = theta[0], theta[1]
    return np.ascontiguousarray(np.column_stack([
        rng.uniform(c / 3, 3 * c, n), rng.uniform(e / 3, min(0.5, 3 * e), n), rng.uniform(0.3, 1.0, n),
        rng.uniform(0.0, 1.0, n), rng.uniform(0.05, 1.0, n)]))
