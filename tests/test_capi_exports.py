"""The C-ABI library loads on a machine without a GPU and exports exactly what include/*.h
declares; without a device every compute entry point fails loudly (no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

from covest_b200 import _capi, build, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'covest_b200.h')) as f:
        text = f.read()
    return set(re.findall(r'CVB_API\s+[\w\s\*]+?\b(cvb_\w+)\s*\(', text))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    names = declared_symbols()
    assert len(names) >= 14
    assert names == set(_capi.EXPORTS)
    for name in names:
        assert getattr(lib, name) is not None
    assert lib.cvb_version().startswith(b'covest_b200')


def test_library_is_sm100a_only():
    out = os.popen('cuobjdump -lelf %s 2>/dev/null' % build.LIB_PATH).read()
    if not out:
        pytest.skip('cuobjdump not available')
    archs = set(re.findall(r'sm_\w+', out))
    assert archs == {'sm_100a'}, archs


def test_no_cpu_fallback_without_a_device():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('a CUDA device is present')
    except ImportError:
        pass
    with pytest.raises(engine.DeviceError) as err:
        engine.LikelihoodContext(0, 21, 100, 8, [1, 2, 3], [5., 4., 3.], 0, None,
                                 ((.01, None), (0, .5)), [1.0] * 8)
    assert 'no CPU path' in str(err.value)
    from covest_b200.models import BasicModel
    with pytest.raises(engine.DeviceError):
        BasicModel(21, 100, {1: 5, 2: 3}, 0, max_error=8).compute_loglikelihood(10, .05)
    lib = _capi.load()
    assert lib.cvb_loglik_batch(None, 1, None, None, None) < 0
    out = ctypes.c_double()
    assert lib.cvb_fp64_peak(None, 0, 1, ctypes.byref(out)) < 0


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'covest_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                if re.search(r'^\s*(from|import)\s+oracle|oracle/|libcovest_oracle|host_math', text, re.M):
                    bad.append(fn)
    # cvpoint.h / cvmath.h mention tests/host_math in comments only; nothing links or imports it
    assert [b for b in bad if b.endswith('.py')] == []
