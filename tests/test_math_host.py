"""Scalar building blocks of the CUDA kernels (covest_b200/csrc/cvmath.h), compiled for the host
and checked against the container's libm / long double.  Test infrastructure only."""
import ctypes
import math
import os
import subprocess

import pytest

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_math')
_SRC = os.path.join(_HERE, 'math_probe.cpp')
_LIB = os.path.join(_HERE, 'libmath_probe.so')
_HDR = os.path.join(os.path.dirname(os.path.dirname(_HERE)), 'covest_b200', 'csrc', 'cvmath.h')


@pytest.fixture(scope='module')
def lib():
    if (not os.path.exists(_LIB)) or any(os.path.getmtime(p) > os.path.getmtime(_LIB) for p in (_SRC, _HDR)):
        subprocess.check_call(['g++', '-O2', '-std=gnu++17', '-ffp-contract=off', '-fPIC', '-shared', _SRC,
                               '-o', _LIB])
    L = ctypes.CDLL(_LIB)
    d = ctypes.c_double
    L.probe_exp_libm.restype = d
    L.probe_exp_libm.argtypes = [d]
    L.probe_one_minus_exp_neg.restype = d
    L.probe_one_minus_exp_neg.argtypes = [d]
    L.probe_pow_uint.restype = d
    L.probe_pow_uint.argtypes = [d, ctypes.c_int]
    L.probe_log_dd.restype = None
    L.probe_log_dd.argtypes = [d, ctypes.POINTER(d), ctypes.POINTER(d)]
    L.probe_exp_mismatches.restype = ctypes.c_long
    L.probe_exp_mismatches.argtypes = [d, d, ctypes.c_long, ctypes.c_int, ctypes.c_ulonglong]
    L.probe_log_tab_worst.restype = d
    L.probe_log_tab_worst.argtypes = [d, d, ctypes.c_long, ctypes.c_ulonglong]
    L.probe_log_tab.restype = d
    L.probe_log_tab.argtypes = [d]
    return L


def test_exp_emulation_is_bit_identical_to_libm(lib):
    """`1.0 - exp(-l)` (models.py:87, :221) needs the rounding of the libm that ran the reference."""
    assert lib.probe_exp_mismatches(-60.0, -1e-12, 400000, 1, 7) == 0
    assert lib.probe_exp_mismatches(-40.0, 0.0, 400000, 0, 9) == 0
    for x in (0.0, -1e-300, -2.0 ** -54, -2.0 ** -53, -1.0, -37.9, -38.0):
        assert lib.probe_exp_libm(x) == math.exp(x)
    assert lib.probe_one_minus_exp_neg(50.0) == 1.0
    assert lib.probe_one_minus_exp_neg(1e-9) == 1.0 - math.exp(-1e-9)


def test_integer_powers_are_correctly_rounded(lib):
    """`(1.0 - err) ** (k - s)`, `err ** s` (models.py:76-78): libm's pow is correctly rounded for
    all but ~1 in 1000 arguments; the kernel's integer power always is."""
    import random

    import mpmath
    mpmath.mp.prec = 400
    rnd = random.Random(3)
    off_libm = 0
    for _ in range(3000):
        x = rnd.uniform(0.0, 1.0)
        n = rnd.randrange(0, 40)
        exact = float(mpmath.mpf(x) ** n)
        assert lib.probe_pow_uint(x, n) == exact
        off_libm += (x ** n != exact)
    assert off_libm <= 30


def test_double_double_log(lib):
    import mpmath
    mpmath.mp.prec = 200
    hi, lo = ctypes.c_double(), ctypes.c_double()
    for x in (1e-300, 3e-8, 0.1, 0.999999, 1.0, 1.5, 2.0, 117.17, 5740.0, 1e10):
        lib.probe_log_dd(x, ctypes.byref(hi), ctypes.byref(lo))
        want = mpmath.log(mpmath.mpf(x))
        got = mpmath.mpf(hi.value) + mpmath.mpf(lo.value)
        assert abs(got - want) <= 2e-18 * max(1.0, abs(want))  # cvmath.h: ~1e-18 |log x| absolute


def test_table_log_of_the_gemm_epilogue(lib):
    """safe_log of a bin probability in the factored path: 2e-16 (1 + |log x|) absolute."""
    assert lib.probe_log_tab_worst(1e-307, 1e-290, 200000, 1) < 3e-16
    assert lib.probe_log_tab_worst(1e-30, 1.0, 400000, 2) < 3e-16
    assert lib.probe_log_tab_worst(0.5, 2.0, 400000, 3) < 3e-16
    assert lib.probe_log_tab_worst(1.0, 1e30, 200000, 4) < 3e-16
    assert lib.probe_log_tab(0.0) == -math.inf
    assert lib.probe_log_tab(1.0) == 0.0 or abs(lib.probe_log_tab(1.0)) < 1e-16
    assert math.isnan(lib.probe_log_tab(math.nan))
    assert lib.probe_log_tab(math.inf) == math.inf
    assert abs(lib.probe_log_tab(5e-324) - math.log(5e-324)) < 1e-12
