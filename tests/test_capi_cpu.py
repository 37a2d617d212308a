"""CPU-only checks of the C-ABI library: it loads, exports every declared symbol, fails loudly without a
GPU, and its host-side quadrature nodes equal the oracle's independent restatement."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from ital_b200 import _capi
from ital_b200.build import build_library
from oracle import orthant


@pytest.fixture(scope='module')
def lib():
    build_library()
    return _capi.load()


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'ital_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ital_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
        assert name in _capi.SIGNATURES, 'binding missing for ' + name
    assert set(_capi.SIGNATURES) == set(names)
    assert lib.ital_version() >= 100


def test_no_gpu_is_a_loud_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from ital_b200 import ITAL
    with pytest.raises(_capi.ItalError, match='no CPU fallback'):
        ITAL(np.random.default_rng(0).uniform(size=(8, 3)), length_scale=1.0)


@pytest.mark.parametrize('t', [1, 2, 3, 4])
def test_snq_nodes_match_oracle(lib, t):
    rng = np.random.default_rng(t)
    for trial in range(4):
        A = rng.normal(size=(t, t + 1))
        C = A @ A.T / (t + 1) * rng.uniform(0.2, 1.0)
        L = np.linalg.cholesky(C)
        m = rng.normal(size=t) * (0.1 if trial < 2 else 1.5)
        eta_o, w_o, orth_o = orthant.snq_nodes(m, L)
        order = np.argsort(orth_o, kind='stable')
        n = lib.ital_snq_nodes(t, _capi.dptr(m), _capi.dptr(np.ascontiguousarray(L)), None, None, None, None)
        assert n == len(w_o) == (2 * lib.ital_snq_order(t)) ** t
        eta = np.zeros((t, n))
        w = np.zeros(n)
        orth = np.zeros(n, dtype=np.int32)
        masses = np.zeros(1 << t)
        Lc = np.ascontiguousarray(L)
        assert lib.ital_snq_nodes(t, _capi.dptr(m), _capi.dptr(Lc), _capi.dptr(eta), _capi.dptr(w),
                                  orth.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _capi.dptr(masses)) == n
        assert np.array_equal(orth, orth_o[order])
        np.testing.assert_allclose(eta.T, eta_o[order], rtol=0, atol=2e-13)
        np.testing.assert_allclose(w, w_o[order], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(masses, orthant.base_masses(w_o, orth_o, t), rtol=1e-12)
        assert abs(masses.sum() - 1.0) < {1: 1e-11, 2: 1e-8, 3: 1e-6, 4: 1e-1}[t]   # t >= 4: coarse panels, see DESIGN.md (batches > 4)


def test_snq_order_matches_oracle(lib):
    for t in range(1, 11):
        assert lib.ital_snq_order(t) == orthant.snq_order(t)
