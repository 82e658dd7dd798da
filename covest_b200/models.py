"""BasicModel / RepeatsModel with the reference's interface (covest/models.py:17-259), evaluated on
a B200 through libcovest_b200.so.

Same constructor arguments, attributes and return types as the reference classes; what changes is
where the arithmetic runs: compute_probabilities / compute_loglikelihood /
compute_loglikelihood_multi hand their arguments to the device (one batched launch), and two batch
entry points are added -- loglikelihood_batch and probabilities_batch -- that the estimator and the
grid search use so that whole candidate sets are one launch.  No likelihood arithmetic is done in
Python; without the compiled library or a CUDA device the evaluators raise.
"""
import inspect
import sys
import warnings

import numpy as np

from . import _capi, constants
from .engine import LikelihoodContext

MODEL_CLASS_SUFFIX = 'Model'


def _float_comb(k):
    """C(k,s) * 3**s as models.py:25 computes it: scipy's floating-point comb (for some (k, s) it
    is an ulp away from the integer), times the Python integer 3**s."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from scipy.special import comb
    return [comb(k, s) * (3 ** s) for s in range(k + 1)]


class BasicModel:
    """Error classes s = 0..max_error-1 of a k-mer, each a zero-truncated Poisson
    (reference: covest/models.py:17-170)."""

    params = ('coverage', 'error_rate')
    _kind = _capi.MODEL_BASIC

    def __init__(self, k, r, hist, tail, max_error=None, max_cov=None, *args, **kwargs):
        self.repeats = False
        self.k = k
        self.r = r
        self.bounds = ((0.01, max_cov), (0, 0.5))
        self.defaults = (1, self._default_param(1))
        self.comb = _float_comb(k)
        self.hist = hist
        self.tail = tail
        self.max_error = k + 1 if max_error is None else min(k + 1, max_error)
        self.threshold = None
        self._ctx = None
        self._ctx_key = None

    # -- names, bounds (models.py:33-69) ----------------------------------------------------
    @classmethod
    def short_name(cls):
        name = cls.__name__
        if name.endswith(MODEL_CLASS_SUFFIX):
            name = name[:-len(MODEL_CLASS_SUFFIX)]
        return name.lower()

    @property
    def param_count(self):
        return len(self.params)

    def _default_param(self, i, default=None):
        lo, hi = self.bounds[i]
        if lo is None or hi is None:
            return default
        return (lo + hi) / 2

    def check_bounds(self, args):
        for arg, (lo, hi) in zip(args, self.bounds):
            if arg is None:
                continue
            if arg == float('NaN'):  # never true; kept for behavioural parity (models.py:53)
                return False
            if (lo is not None and arg < lo) or (hi is not None and arg > hi):
                return False
        return True

    def fit_to_bounds(self, args):
        out = list(args)
        for i, (arg, (lo, hi)) in enumerate(zip(out, self.bounds)):
            if arg is None:
                continue
            if lo is not None and arg < lo:
                out[i] = lo
            elif hi is not None and arg > hi:
                out[i] = hi
        return out

    def correct_c(self, c):
        return c * (self.r - self.k + 1) / self.r

    def _get_lambda_s(self, c, err):
        return [c * (3 ** -s) * (1.0 - err) ** (self.k - s) * err ** s
                for s in range(self.max_error)]

    # -- device context ---------------------------------------------------------------------
    def _state_key(self):
        # the reference reads self.hist on every call (models.py:92, :235): counts changed in place
        # must reach the device copy, so the items themselves are part of the key (a few hundred to
        # a few thousand bins)
        return (hash(tuple(self.hist.items())), self.tail, self.max_error, self.bounds,
                self.threshold, self.k, self.r)

    @property
    def device_context(self):
        """The histogram on the device (created on first use; rebuilt if the model's state
        changed)."""
        key = self._state_key()
        if self._ctx is None or key != self._ctx_key:
            if self._ctx is not None:
                self._ctx.close()
            if not self.hist:
                raise ValueError('empty histogram')
            keys = list(self.hist.keys())
            counts = [float(self.hist[j]) for j in keys]
            self._ctx = LikelihoodContext(self._kind, self.k, self.r, self.max_error, keys, counts,
                                          float(self.tail), self.threshold, self.bounds, self.comb)
            self._ctx_key = key
        return self._ctx

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_ctx'] = None
        state['_ctx_key'] = None
        return state

    def _rows(self, args_list):
        n = self.param_count
        rows = np.empty((len(args_list), n), dtype=np.float64)
        for i, args in enumerate(args_list):
            rows[i] = [float(a) for a in list(args)[:n]]
        return rows

    # -- batched evaluators (new) -----------------------------------------------------------
    def loglikelihood_batch(self, points):
        """Log-likelihood of every row of `points` (n x param_count), clipped to the bounds like
        compute_loglikelihood; one launch."""
        return self.device_context.loglik(points)

    def probabilities_batch(self, points, clip=False):
        """compute_probabilities for every row: (n, len(hist)) array, columns in hist key order."""
        return self.device_context.probs(points, clip=clip)

    # -- reference interface (models.py:81-117) ---------------------------------------------
    def compute_probabilities(self, c, err, *rest):
        row = [c, err] + list(rest)[:self.param_count - 2]
        p = self.device_context.probs(np.array([row], dtype=np.float64), clip=False)[0]
        return {j: float(v) for j, v in zip(self.hist.keys(), p)}

    def compute_loglikelihood(self, *args):
        if not self.hist:  # models.py:100-107 on an empty dict: no bins, no mass
            if self.repeats:
                raise ValueError('max() arg is an empty sequence')
            return 0.0
        return float(self.device_context.loglik(self._rows([args]))[0])

    def compute_loglikelihood_multi(self, args_list, thread_count=constants.DEFAULT_THREAD_COUNT):
        """{tuple(args): loglikelihood}; `thread_count` is accepted for compatibility and
        ignored -- the batch is one device launch (the reference forks a Pool, models.py:113)."""
        args_list = [tuple(a) for a in args_list]
        if not args_list:
            return {}
        ll = self.device_context.loglik(self._rows(args_list))
        return {args: float(v) for args, v in zip(args_list, ll)}

    def plot_probs(self, est, guess, orig, cumulative=False, log_scale=True):
        """Histogram vs fitted probabilities (models.py:119-170).  Needs matplotlib."""
        import matplotlib.pyplot as plt
        keys = sorted(self.hist)
        total = sum(self.hist.values())
        series = [('data', [self.hist[j] / total for j in keys])]
        for label, args in (('estimate', est), ('guess', guess), ('original', orig)):
            if args is None or any(a is None for a in args):
                continue
            p = self.compute_probabilities(*args)
            series.append((label, [p[j] for j in keys]))
        for label, ys in series:
            if cumulative:
                ys = np.cumsum(ys)
            plt.plot(keys, ys, label=label)
        if log_scale:
            plt.yscale('log')
        plt.legend()
        plt.show()


class RepeatsModel(BasicModel):
    """Copy-number mixture over o = 1, 2, 3.. copies on top of the error classes
    (reference: covest/models.py:173-242)."""

    params = BasicModel.params + ('q1', 'q2', 'q')
    _kind = _capi.MODEL_REPEATS

    def __init__(self, k, r, hist, tail, max_error=None, max_cov=None, threshold=1e-8,
                 min_single_copy_ratio=0.3, *args, **kwargs):
        # like the reference (models.py:177) max_cov is not forwarded: coverage has no upper bound
        super().__init__(k, r, hist, tail, max_error=max_error)
        self.repeats = True
        self.bounds = self.bounds + ((min_single_copy_ratio, 1), (0, 1), (0, 1))
        self.defaults = self.defaults + tuple(self._default_param(i, default=0.5)
                                              for i in range(2, 5))
        self.threshold = threshold

    def get_hist_threshold(self, b_o, threshold):
        top = max(self.hist)
        if threshold is not None:
            for o in range(1, top):
                if b_o(o) <= threshold:
                    return o
        return top

    @staticmethod
    def get_b_o(q1, q2, q):
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

        return b_o

    # noinspection PyMethodOverriding
    def compute_probabilities(self, c, err, q1, q2, q, *_):
        return super().compute_probabilities(c, err, q1, q2, q)


models = {
    cls.short_name(): cls for _, cls in inspect.getmembers(
        sys.modules[__name__],
        predicate=lambda x: inspect.isclass(x) and x.__name__.endswith(MODEL_CLASS_SUFFIX))
}


def select_model(m):
    """Full name or any prefix of it (models.py:252-259)."""
    if m in models:
        return models[m]
    for name, model in models.items():
        if name.startswith(m):
            return model
    raise ValueError('Not such model: {}.'.format(m))
