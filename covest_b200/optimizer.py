"""Lock-step multi-start optimiser: every start advances one iteration per pair of device launches.

The reference refines each start with its own scipy L-BFGS-B run in a forked process
(covest/covest.py:33-39, :63-78): ~190 objective calls per start, one point at a time, on forward
differences of 1e-8 that stop ~1e-4 (relative) short of the optimum.
Here all starts move together.  One iteration is

  launch 1   the central-difference stencils of every running start (1 + 2m + 2m(m-1) points for m
             free coordinates: 51 for the repeats model) -> gradient and Hessian in the scaled
             coordinates x_i / h_i
  host       per start a projected, eigenvalue-modified Newton direction (m <= 5: microseconds)
  launch 2   a ladder of step lengths along it for every start; the best strictly decreasing one is
             taken

so 16 starts cost ~2 launches x ~20 iterations instead of ~3 800 sequential evaluations, and the
iteration ends at the optimum itself (it is the Newton polish of CoverageEstimator.polish, run from
the start).  `evaluate(points)` is any batched objective (CoverageEstimator.likelihood_batch).
"""
import numpy as np

STEP_LADDER = (1.0, 0.5, 0.25, 1.0 / 16, 1.0 / 64, 1.0 / 256, 1.0 / 2048)


def _stencil(m):
    """Offsets (in units of h) of the central-difference stencil over m coordinates: the centre,
    +-e_a, and (+-e_a +-e_b) for a < b."""
    rows = [np.zeros(m)]
    for a in range(m):
        for s in (1.0, -1.0):
            e = np.zeros(m)
            e[a] = s
            rows.append(e)
    pairs = [(a, b) for a in range(m) for b in range(a + 1, m)]
    for a, b in pairs:
        for sa, sb in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
            e = np.zeros(m)
            e[a], e[b] = sa, sb
            rows.append(e)
    return np.array(rows), pairs


class LockstepResult:
    def __init__(self, x, fun, success, nit, nfev):
        self.x, self.fun, self.success, self.nit, self.nfev = x, fun, success, nit, nfev

    def __repr__(self):
        return 'LockstepResult(x=%r, fun=%r, success=%r, nit=%d)' % (self.x.tolist(), self.fun, self.success, self.nit)


def lockstep_minimize(evaluate, starts, bounds, fixed=None, rel_step=1e-4, max_iter=80, xtol=1e-10,
                      ftol=1e-10, max_scaled_step=5000.0):
    """Minimise `evaluate` from every row of `starts` at once.  bounds: [(lo, hi)] with None for an
    open end; fixed: boolean mask of coordinates that do not move.  Returns a list of
    LockstepResult (x, fun, success, nit, nfev), one per start, plus the number of launches.
    A start stops when its step falls below xtol (relative), when no step length decreases the
    objective, or when two consecutive iterations gain less than ftol (relative; scipy's L-BFGS-B
    stops at 2.2e-9) -- the flat valleys of unidentifiable parameters (q2 when q1 is 1)."""
    X = np.array(starts, dtype=np.float64).reshape(len(starts), -1)
    S, n = X.shape
    lo = np.array([-np.inf if b[0] is None else b[0] for b in bounds], dtype=np.float64)
    hi = np.array([np.inf if b[1] is None else b[1] for b in bounds], dtype=np.float64)
    fixed = np.zeros(n, dtype=bool) if fixed is None else np.asarray(fixed, dtype=bool)
    launches = 0
    nfev = np.zeros(S, dtype=np.int64)
    nit = np.zeros(S, dtype=np.int64)
    if S == 0:
        return [], 0
    X = np.minimum(np.maximum(X, lo), hi)
    F = np.asarray(evaluate(X), dtype=np.float64).copy()
    launches += 1
    nfev += 1
    running = np.isfinite(F)
    ok = np.zeros(S, dtype=bool)
    flat_steps = np.zeros(S, dtype=np.int64)
    small_gain = np.zeros(S, dtype=np.int64)
    # a coordinate whose interval is narrower than its stencil cannot be differentiated: it stays
    span = np.where(np.isfinite(hi - lo), hi - lo, np.inf)
    free = ~fixed & (span > 4 * rel_step * np.maximum(np.maximum(np.abs(lo), np.abs(np.where(np.isfinite(hi), hi, 0))), 1e-3))
    fi = np.flatnonzero(free)
    m = len(fi)
    if m == 0:
        return [LockstepResult(X[s], float(F[s]), bool(running[s]), 0, 1) for s in range(S)], launches
    E, pairs = _stencil(m)
    P = len(E)
    ladder = np.array(STEP_LADDER)
    for _ in range(max_iter):
        act = np.flatnonzero(running)
        if len(act) == 0:
            break
        Xa = X[act]
        H = rel_step * np.maximum(np.abs(Xa[:, fi]), 1e-3)                     # (A, m)
        C = Xa.copy()
        C[:, fi] = np.minimum(np.maximum(Xa[:, fi], lo[fi] + H), hi[fi] - H)   # stencils stay inside the bounds
        pts = np.repeat(C[:, None, :], P, axis=1)                              # (A, P, n)
        pts[:, :, fi] += E[None, :, :] * H[:, None, :]
        vals = np.asarray(evaluate(pts.reshape(-1, n)), dtype=np.float64).reshape(len(act), P)
        launches += 1
        nfev[act] += P
        nit[act] += 1
        f0 = vals[:, 0]
        fp, fm = vals[:, 1:1 + 2 * m:2], vals[:, 2:2 + 2 * m:2]
        gs = 0.5 * (fp - fm)                                                   # gradient, scaled by h
        Hs = np.zeros((len(act), m, m))
        ar = np.arange(m)
        Hs[:, ar, ar] = fp - 2 * f0[:, None] + fm
        base = 1 + 2 * m
        for q, (a, b) in enumerate(pairs):
            blk = vals[:, base + 4 * q: base + 4 * q + 4]
            v = 0.25 * (blk[:, 0] - blk[:, 1] - blk[:, 2] + blk[:, 3])
            Hs[:, a, b] = v
            Hs[:, b, a] = v
        good = np.all(np.isfinite(vals), axis=1)
        # coordinates held at a bound: the descent direction points out of the box
        at_lo = (Xa[:, fi] <= lo[fi] + H) & (gs > 0)
        at_hi = (Xa[:, fi] >= hi[fi] - H) & (gs < 0)
        pinned = at_lo | at_hi
        D = np.zeros((len(act), m))
        for r in np.flatnonzero(good):
            keep = np.flatnonzero(~pinned[r])
            if len(keep) == 0:
                continue
            w, V = np.linalg.eigh(Hs[r][np.ix_(keep, keep)])
            floor = max(1e-10 * np.max(np.abs(w)), 1e-300)
            w = np.maximum(np.abs(w), floor)       # a descent direction also where the surface is not convex
            d = -V @ ((V.T @ gs[r, keep]) / w)
            big = np.max(np.abs(d))
            if big > max_scaled_step:
                d *= max_scaled_step / big
            D[r, keep] = d
        # gradient fallback where the stencil left the objective's finite domain
        for r in np.flatnonzero(~good):
            g = np.where(np.isfinite(gs[r]), gs[r], 0.0)
            D[r] = -g / max(np.max(np.abs(g)), 1e-300)
        step = D * H                                                            # back to x
        base_x = Xa.copy()
        base_x[:, fi] = np.where(at_lo, lo[fi], np.where(at_hi, hi[fi], Xa[:, fi]))
        cand = np.repeat(base_x[:, None, :], len(ladder), axis=1)              # (A, L, n)
        cand[:, :, fi] += ladder[None, :, None] * step[:, None, :]
        cand = np.minimum(np.maximum(cand, lo), hi)
        cv = np.asarray(evaluate(cand.reshape(-1, n)), dtype=np.float64).reshape(len(act), len(ladder))
        launches += 1
        nfev[act] += len(ladder)
        cv = np.where(np.isfinite(cv), cv, np.inf)
        pick = np.argmin(cv, axis=1)
        best = cv[np.arange(len(act)), pick]
        for r, s in enumerate(act):
            if best[r] < F[s]:
                new = cand[r, pick[r]]
                moved = np.max(np.abs(new - X[s]) / np.maximum(np.abs(X[s]), 1e-12))
                # long steps for no gain: a valley, not the quadratic end game (where steps shrink)
                valley = moved >= 1e-4 and F[s] - best[r] <= ftol * max(abs(F[s]), 1.0)
                small_gain[s] = small_gain[s] + 1 if valley else 0
                X[s] = new
                F[s] = best[r]
                if moved < xtol or small_gain[s] >= 2:
                    running[s] = False
                    ok[s] = True
            elif good[r] and flat_steps[s] < 4 and np.isfinite(cv[r, 0]) and np.max(np.abs(D[r])) < 10.0:
                # The objective is flat to its last bits (one ulp of 4e6 is 5e-10) but the gradient
                # still resolves the optimum: take the small full Newton step regardless -- the
                # polish of SURVEY.md section 7.3 item 3.
                new = cand[r, 0]
                moved = np.max(np.abs(new - X[s]) / np.maximum(np.abs(X[s]), 1e-12))
                X[s] = new
                F[s] = cv[r, 0]
                flat_steps[s] += 1
                if moved < 1e-9:
                    running[s] = False
                    ok[s] = True
            else:  # no step length decreases the objective any more: the optimum at this resolution
                running[s] = False
                ok[s] = bool(good[r])
    # a start that is still descending after max_iter reports what it reached, unsuccessfully
    return [LockstepResult(X[s].copy(), float(F[s]), bool(ok[s]), int(nit[s]), int(nfev[s]))
            for s in range(S)], launches
