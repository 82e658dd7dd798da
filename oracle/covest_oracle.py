"""ctypes front end of the CPU oracle (oracle/covest_oracle.c) plus a small pure-Python
restatement that drives the reference's own compiled C module when oracle/_ref/ has it.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (covest_b200/) never imports it.

Reference lines are relative to /root/reference.
"""
import ctypes
import math
import os
import subprocess
import sys
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libcovest_oracle.so')
_REF_DIR = os.path.join(_HERE, '_ref')

BASIC, REPEATS = 0, 1
FAITHFUL, LADDER = 0, 1  # one truncated_poisson call per term / bit-identical linear pass


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, the reference's own
    C module into oracle/_ref/)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, 'covest_oracle.c'))):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []) + ["libcovest_oracle.so"])
    if os.path.exists('/root/reference/c_src/covest_poissonmodule.c'):
        subprocess.check_call(['make', '-C', _HERE, 'ref'])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.cvo_truncated_poisson.restype = ctypes.c_double
        L.cvo_truncated_poisson.argtypes = [ctypes.c_double, ctypes.c_int]
        L.cvo_probs.restype = ctypes.c_int
        L.cvo_probs.argtypes = [ctypes.c_int] * 5 + [ip, dp, ctypes.c_double, dp, ctypes.c_int, dp]
        L.cvo_loglik_batch.restype = ctypes.c_int
        L.cvo_loglik_batch.argtypes = [ctypes.c_int] * 5 + [ip, dp, ctypes.c_double, dp,
                                                            ctypes.c_double, dp, ctypes.c_long, dp,
                                                            ctypes.c_int, ctypes.c_int, dp]
        _lib = L
    return _lib


def reference_comb(k):
    """comb[s] = C(k,s) * 3**s exactly as models.py:25 computes it (scipy's floating comb,
    which is NOT always the exact integer: e.g. k=31, s=14)."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from scipy.special import comb
    # the entries stay numpy.float64 scalars, as in the reference: that is what makes Python's
    # sum() over them take its generic (uncompensated) path in py_probs below
    return [comb(k, s) * (3 ** s) for s in range(k + 1)]


class Model:
    """What BasicModel/RepeatsModel.__init__ store (models.py:19-31, :175-183)."""

    def __init__(self, kind, k, r, hist, tail=0, max_error=None, max_cov=None, threshold=1e-8,
                 min_single_copy_ratio=0.3):  # noqa: D107
        self.kind = REPEATS if kind in (REPEATS, 'repeats', 'repeat', 'r') else BASIC
        self.k, self.r = int(k), int(r)
        self.hist = dict(hist)
        self.tail = tail
        self.max_error = k + 1 if max_error is None else min(k + 1, max_error)
        self.comb = reference_comb(k)
        if self.kind == REPEATS:
            # models.py:177 does not forward max_cov
            self.bounds = ((0.01, None), (0, 0.5), (min_single_copy_ratio, 1), (0, 1), (0, 1))
        else:
            self.bounds = ((0.01, max_cov), (0, 0.5))
        self.threshold = threshold
        self.bin_j = np.ascontiguousarray(list(self.hist.keys()), dtype=np.int32)
        self.bin_h = np.ascontiguousarray([float(v) for v in self.hist.values()], dtype=np.float64)
        self._comb = np.ascontiguousarray(self.comb[:self.max_error], dtype=np.float64)
        b = []
        for lo, hi in self.bounds:
            b += [math.nan if lo is None else float(lo), math.nan if hi is None else float(hi)]
        self._bounds = np.ascontiguousarray(b, dtype=np.float64)

    @property
    def n_params(self):
        return 5 if self.kind == REPEATS else 2

    def _thr(self):
        return math.nan if self.threshold is None else float(self.threshold)

    def probs(self, params, mode=LADDER):
        """compute_probabilities(*params) as an array aligned with list(hist)."""
        par = np.zeros(5, dtype=np.float64)
        par[:self.n_params] = [float(x) for x in params[:self.n_params]]
        out = np.empty(len(self.bin_j), dtype=np.float64)
        rc = lib().cvo_probs(self.kind, self.k, self.r, self.max_error, len(self.bin_j),
                             _ip(self.bin_j), _dp(self._comb), self._thr(), _dp(par), mode, _dp(out))
        if rc:
            raise RuntimeError('oracle cvo_probs failed: %d' % rc)
        return out

    def loglik_batch(self, points, mode=LADDER, threads=1):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.n_params)
        out = np.empty(len(pts), dtype=np.float64)
        rc = lib().cvo_loglik_batch(self.kind, self.k, self.r, self.max_error, len(self.bin_j),
                                    _ip(self.bin_j), _dp(self.bin_h), float(self.tail),
                                    _dp(self._comb), self._thr(), _dp(self._bounds), len(pts),
                                    _dp(pts), mode, threads, _dp(out))
        if rc:
            raise RuntimeError('oracle cvo_loglik_batch failed: %d' % rc)
        return out

    def loglik(self, *params, mode=LADDER):
        return float(self.loglik_batch([list(params[:self.n_params])], mode=mode)[0])


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def truncated_poisson(l, j):
    return lib().cvo_truncated_poisson(float(l), int(j))


# ---------------------------------------------------------------------------------------------
# The reference's own compiled C module (oracle/_ref/covest_poisson*.so, built by `make ref`
# from /root/reference/c_src/covest_poissonmodule.c) driven by a pure-Python restatement of
# models.py:81-107 / :211-242.  Slow (one C call per term) -- small cases and the
# `--impl reference` CPU timing only.
# ---------------------------------------------------------------------------------------------
def ref_module():
    """The compiled reference module or None."""
    if not os.path.isdir(_REF_DIR):
        return None
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    try:
        import covest_poisson
        return covest_poisson
    except ImportError:
        return None


def _guard_zero_rate(tp):
    """`truncated_poisson(0, j)` is undefined behaviour in the reference (c:15-17 passes an int
    through "d" varargs): it returns whatever the xmm register held, sometimes NaN, always times a
    weight that is exactly 0.  This wrapper returns the 0 the source intends, so that the CPU arm
    does not depend on register garbage."""
    def guarded(l, j):
        return tp(l, j) if l else 0.0
    return guarded


def py_probs(model, params, tp):
    """models.py:81-98 (basic) / :211-242 (repeats) with `tp` as truncated_poisson."""
    k, r, S = model.k, model.r, model.max_error
    c, err = params[0], params[1]
    ck = c * (r - k + 1) / r
    l_s = [ck * (3 ** -s) * (1.0 - err) ** (k - s) * err ** s for s in range(S)]
    comb = model.comb
    if not all(l_s):  # a zero rate (err = 0): see _guard_zero_rate; no overhead on ordinary points
        tp = _guard_zero_rate(tp)
    if model.kind == BASIC:
        n_s = [comb[s] * (1.0 - math.exp(-l_s[s])) for s in range(S)]
        tot = sum(n_s)
        if tot == 0:
            tot = 1
        a_s = [n / tot for n in n_s]
        return {j: sum(a_s[s] * tp(l_s[s], j) for s in range(S)) for j in model.hist}
    q1, q2, q = params[2], params[3], params[4]
    two = (1 - q1) * q2
    many = (1 - q1) * (1 - q2) * q

    def b_o(o):
        if o == 0:
            return 0
        if o == 1:
            return q1
        if o == 2:
            return two
        return many * (1 - q) ** (o - 3)

    top = max(model.hist)
    cut = top
    if model.threshold is not None:
        for o in range(1, top):
            if b_o(o) <= model.threshold:
                cut = o
                break
    a_os = {}
    for o in range(1, cut):
        n = [comb[s] * (1.0 - math.exp(o * -l_s[s])) for s in range(S)]
        tot = sum(n)
        if tot == 0:
            tot = 1
        a_os[o] = [x / tot for x in n]
    return {
        j: sum(b_o(o) * sum(a_os[o][s] * tp(o * l_s[s], j) for s in range(S)) for o in range(1, cut))
        for j in model.hist
    }


def py_loglik(model, params, tp):
    """models.py:100-107."""
    args = list(params[:model.n_params])
    for i, (lo, hi) in enumerate(model.bounds):
        if lo is not None and args[i] < lo:
            args[i] = lo
        elif hi is not None and args[i] > hi:
            args[i] = hi
    p = py_probs(model, args, tp)
    mass = min(1, math.fsum(p.values()))
    slog = lambda x: -math.inf if x <= 0 else math.log(x)  # noqa: E731
    tail = model.tail * slog(1 - mass) if mass < 1 else 0
    return float(sum(h * slog(p[j]) for j, h in model.hist.items() if h)) + tail


def _pool_eval(job):
    model, params = job
    return py_loglik(model, params, ref_module().truncated_poisson)


def ref_pool(processes):
    """A fork pool for ref_loglik_batch (the reference forks one per call, models.py:113; a caller
    that times the evaluation creates it once, outside the timed region)."""
    import multiprocessing
    return multiprocessing.get_context('fork').Pool(processes)


def ref_loglik_batch(model, points, processes=1, pool=None):
    """Log-likelihoods through the reference's compiled module; a fork pool over points is the
    reference's own parallelism (models.py:109-117)."""
    if ref_module() is None:
        raise RuntimeError('oracle/_ref is not built (make -C oracle ref needs /root/reference)')
    jobs = [(model, list(map(float, p))) for p in points]
    if pool is not None:
        return np.array(pool.map(_pool_eval, jobs, chunksize=1))
    if processes <= 1:
        return np.array([_pool_eval(j) for j in jobs])
    with ref_pool(processes) as own:
        return np.array(own.map(_pool_eval, jobs))
