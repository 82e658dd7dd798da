"""ctypes front end of the test-only host emulation of the kernel phases
(tests/host_math/emulate.cpp).  Test infrastructure: never imported by covest_b200/."""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_math')
_SRC = os.path.join(_HERE, 'emulate.cpp')
_LIB = os.path.join(_HERE, 'libcv_emulate.so')
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'covest_b200', 'csrc')

_lib = None


def _stale():
    if not os.path.exists(_LIB):
        return True
    t = os.path.getmtime(_LIB)
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith('.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            subprocess.check_call(['g++', '-O2', '-std=gnu++17', '-ffp-contract=off', '-fPIC',
                                   '-shared', _SRC, '-o', _LIB])
        L = ctypes.CDLL(_LIB)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.emu_loglik_batch.restype = ctypes.c_int
        L.emu_loglik_batch.argtypes = [ctypes.c_int] * 5 + [ip, dp, ctypes.c_double, dp, dp,
                                                            ctypes.c_double, dp, ctypes.c_long, dp,
                                                            ctypes.c_int, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def loglik_batch(model, points, clip=True, want_probs=False):
    """`model` is an oracle.covest_oracle.Model (only its stored description is used)."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, model.n_params)
    S = model.max_error
    comb = np.ascontiguousarray(model.comb[:S], dtype=np.float64)
    pow3 = np.ascontiguousarray([1.0 if s == 0 else float(3 ** -s) for s in range(S)])
    out = np.empty(len(pts))
    probs = np.zeros((len(pts), len(model.bin_j))) if want_probs else None
    thr = math.nan if model.threshold is None else float(model.threshold)
    rc = lib().emu_loglik_batch(model.kind, model.k, model.r, S, len(model.bin_j),
                                model.bin_j.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                _dp(model.bin_h), float(model.tail), _dp(comb), _dp(pow3), thr,
                                _dp(model._bounds), len(pts), _dp(pts), int(clip), _dp(out),
                                _dp(probs) if want_probs else None)
    if rc:
        raise RuntimeError('emulation failed: %d' % rc)
    return (out, probs) if want_probs else out


BAND_LL = -1.0e270  # CV_BAND_LL of csrc/cvmodel.h


def marked(values):
    """Points the epilogue marks for the term-by-term re-evaluation (a bin with a count has a
    probability in the subnormal range; csrc/cvmodel.h CV_PSCALE).  That re-evaluation is a CUDA
    kernel (csrc/faithful.cu) and is checked on the GPU; the host emulation stops at the mark."""
    v = np.asarray(values, dtype=float)
    return np.isfinite(v) & (v < BAND_LL)
