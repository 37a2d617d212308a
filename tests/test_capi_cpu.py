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


@pytest.mark.parametrize('t', [1, 2, 3, 4, 5, 6, 7])
def test_snq_nodes_match_oracle(lib, t):
    """Host node generation (csrc/snq_host.h) against the oracle's independent restatement (oracle/orthant.py), node for
    node: tensor rule for t <= 5 base variables, sequential-conditioning lattice inside every orthant from t = 6 on."""
    rng = np.random.default_rng(t)
    for trial in range(4 if t <= 3 else 1):
        A = rng.normal(size=(t, t + 1))
        C = A @ A.T / (t + 1) * rng.uniform(0.2, 1.0)
        L = np.linalg.cholesky(C)
        m = rng.normal(size=t) * (0.1 if trial < 2 else 1.5)
        eta_o, w_o, orth_o = orthant.snq_nodes(m, L)
        order = np.argsort(orth_o, kind='stable')
        cap = lib.ital_snq_nodes(t, _capi.dptr(m), _capi.dptr(np.ascontiguousarray(L)), None, None, None, None)
        assert cap == ((2 * lib.ital_snq_order(t)) ** t if t <= 5 else orthant.SNQ_SC_N + (orthant.SNQ_SC_MIN << t)) >= len(w_o)
        eta_buf = np.zeros(t * cap)
        w = np.zeros(cap)
        orth = np.zeros(cap, dtype=np.int32)
        masses = np.zeros(1 << t)
        Lc = np.ascontiguousarray(L)
        n = lib.ital_snq_nodes(t, _capi.dptr(m), _capi.dptr(Lc), _capi.dptr(eta_buf), _capi.dptr(w),
                               orth.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _capi.dptr(masses))
        assert n == len(w_o)        # both drop the nodes lighter than 1e-13 / share out the lattice nodes alike
        eta, w, orth = eta_buf[:t * n].reshape(t, n), w[:n], orth[:n]
        assert np.array_equal(orth, orth_o[order])
        # (t >= 6: inverse normal CDF of truncated-normal quantiles; two implementations of the inverse, and a
        # quantile at the edge of a half-line is ill-conditioned in eta but carries no weight difference)
        np.testing.assert_allclose(eta.T, eta_o[order], rtol=0, atol=2e-13 if t <= 5 else 1e-8)
        np.testing.assert_allclose(w, w_o[order], rtol=1e-12 if t <= 5 else 1e-9, atol=1e-300)
        np.testing.assert_allclose(masses, orthant.base_masses(w_o, orth_o, t), rtol=1e-12 if t <= 5 else 1e-9, atol=1e-18)
        assert abs(masses.sum() - 1.0) < {1: 1e-11, 2: 1e-8, 3: 1e-6, 4: 1e-6, 5: 1e-5, 6: 1e-12, 7: 1e-12}[t]     # (t >= 6: normalised)


def test_snq_order_matches_oracle(lib):
    for t in range(1, 11):
        assert lib.ital_snq_order(t) == orthant.snq_order(t)


def _general_mi_from_c_arrays(lib, m_b, L, m_c, l_c, s_c, lp, mp, noise):
    """The assembly of k_eval_general restated in numpy on the node sets the C library generates."""
    from scipy.special import ndtr
    t = len(m_b)
    D = t + 1
    sizes = np.zeros(4, dtype=np.int64)
    Lc = np.ascontiguousarray(L)
    i32 = ctypes.POINTER(ctypes.c_int32)
    assert lib.ital_snq_general(t, _capi.dptr(m_b), _capi.dptr(Lc), noise, _capi.i64ptr(sizes),
                                None, None, None, None, None, None) == 0
    N, G, NS, nl = [int(x) for x in sizes]
    eta, w = np.zeros((t, N)), np.zeros(N)
    gb, gm = np.zeros(G + 1, dtype=np.int32), np.zeros(G)
    s0, lut = np.zeros(NS + 1, dtype=np.int32), np.zeros(nl, dtype=np.int32)
    assert lib.ital_snq_general(t, _capi.dptr(m_b), _capi.dptr(Lc), noise, _capi.i64ptr(sizes), _capi.dptr(eta),
                                _capi.dptr(w), gb.ctypes.data_as(i32), _capi.dptr(gm), s0.ctypes.data_as(i32),
                                lut.ctypes.data_as(i32)) == 0
    lut = lut.reshape(1 << D, 1 << D, 3)
    arg = m_c[:, None] + l_c @ eta
    st = np.sqrt(s_c ** 2 + noise)
    cdf = ndtr(arg / s_c[:, None]) * w
    bp = np.exp(-0.5 * ((1 - arg) / st[:, None]) ** 2) * w
    bm = np.exp(-0.5 * ((-1 - arg) / st[:, None]) ** 2) * w
    A = np.stack([cdf[:, gb[g]:gb[g + 1]].sum(axis=1) for g in range(G)], axis=1)
    Bp = np.stack([bp[:, gb[g]:gb[g + 1]].sum(axis=1) for g in range(G)], axis=1)
    Bm = np.stack([bm[:, gb[g]:gb[g + 1]].sum(axis=1) for g in range(G)], axis=1)
    sBp = np.stack([Bp[:, s0[k]:s0[k + 1]].sum(axis=1) for k in range(NS)], axis=1)
    sBm = np.stack([Bm[:, s0[k]:s0[k + 1]].sum(axis=1) for k in range(NS)], axis=1)
    eps = 1e-12
    mi = np.zeros(len(m_c))
    for r in range(1 << D):
        g0, rc = r & ((1 << t) - 1), r >> t
        p_r = np.maximum(A[:, g0] if rc else gm[g0] - A[:, g0], 0)
        inner = -(1 - (1 - lp) ** D) * np.log(p_r + eps)
        for Om in range(1, 1 << D):
            k = bin(Om).count('1')
            g, sidx, flags = lut[r, Om]
            if flags & 2:
                q = np.ones(len(m_c))
            elif flags & 1:
                q = (Bp[:, g] / np.maximum(sBp[:, sidx], 1e-300)) if rc else (Bm[:, g] / np.maximum(sBm[:, sidx], 1e-300))
            else:
                q = A[:, g] if rc else gm[g] - A[:, g]
            q = np.clip(q, 0, 1)
            inner = inner + (1 - lp) ** (D - k) * lp ** k * ((1 - mp) ** k * np.log(q + eps) + (1 - (1 - mp) ** k) * np.log(eps))
        mi += p_r * inner
    return mi


@pytest.mark.parametrize('t,lp,mp', [(1, 0.75, 0.2), (2, 0.25, 0.0), (3, 0.6, 0.1)])
def test_general_feedback_node_sets_match_oracle(lib, t, lp, mp):
    """Host-side conditional node sets + lookup table of the general feedback model against the oracle's literal
    enumeration of relevance and feedback configurations (ital.py:183-224) on random blocks."""
    from oracle.ital_oracle import OracleITAL
    rng = np.random.default_rng(10 * t)
    D = t + 1
    n = 6
    ora = OracleITAL.__new__(OracleITAL)
    ora.label_prob, ora.mistake_prob, ora.label_estimation, ora.noise = lp, mp, 'mean', 1e-6
    A = rng.normal(size=(t, t + 2))
    C = A @ A.T / (t + 2) * 0.5
    L = np.linalg.cholesky(C)
    m_b = rng.normal(size=t) * 0.4
    l_c = rng.normal(size=(n, t)) * 0.3
    s_c = rng.uniform(0.3, 0.8, size=n)
    m_c = rng.normal(size=n) * 0.5
    got = _general_mi_from_c_arrays(lib, m_b, L, m_c, l_c, s_c, lp, mp, 1e-6)
    for i in range(n):
        cbc = L @ l_c[i]
        var_c = s_c[i] ** 2 + l_c[i] @ l_c[i]
        want = ora._mi_general(list(range(D)), m_b, C, m_c[i], var_c, cbc)
        assert abs(got[i] - want) <= 1e-4 * abs(want) + 1e-6, (i, got[i], want)


def test_library_is_sm100a_code_with_bulk_copies_and_dependent_launch():
    """The built library holds sm_100a SASS, the row ring uses the bulk-copy engine (UBLKCP + mbarrier SYNCS), every
    kernel takes part in programmatic dependent launch (ACQBULK / PREEXIT), and the peer exchange uses system-scope
    release / acquire (B200_PROFILING.md: the SASS mnemonics that prove the native path; no GPU needed)."""
    import shutil
    import subprocess
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(tool):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([tool, '-sass', build_library()], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert set(re.findall(r'arch = (\S+)', sass)) == {'sm_100a'}
    ring = [f for f in re.split(r'\n\s*Function : ', sass)[1:] if 'k_extend_bulk' in f.split('\n', 1)[0]]
    assert len(ring) >= 4
    for f in ring:
        assert 'UBLKCP' in f and 'SYNCS.ARRIVE.TRANS64' in f and 'TRYWAIT' in f
    kernels = re.split(r'\n\s*Function : ', sass)[1:]
    fused = [f for f in kernels if 'k_fetch_fused' in f.split('\n', 1)[0]]
    assert len(fused) == 2          # the persistent fetch kernel (cooperative launch): grid barrier = acquire + L1 invalidate
    for f in fused:
        assert 'LDG.E.STRONG.GPU' in f and 'CCTL.IVALL' in f and 'STRONG.SYS' in f
    assert all('ACQBULK' in f and 'PREEXIT' in f for f in kernels if f not in fused), 'a kernel without pdl_enter()'
    assert 'STRONG.SYS' in sass and 'MEMBAR.SC.SYS' in sass


def test_first_step_table_matches_the_closed_form():
    """h_tab (csrc/ital_kernels.cuh): the polynomial table of H(u) = -sum_r p_r log(p_r + eps), p = Phi(+-u), against
    scipy's erfc / log at random and at interval-boundary arguments; what the reference computes with norm.cdf for a
    single sample (ital/ital.py:364-369, 205-219)."""
    from scipy.special import ndtr
    lib = _capi.load()
    n = int(lib.ital_h_table(None, 0))
    tab = np.zeros(n)
    assert lib.ital_h_table(_capi.dptr(tab), n) == n == 136 * 8
    rng = np.random.default_rng(5)
    u = np.concatenate((rng.uniform(0, 8.5, 20000), np.arange(0, 137) / 16.0, np.arange(1, 137) / 16.0 - 1e-13, [8.5]))
    u = np.minimum(u, 8.5)
    k = np.minimum((u * 16).astype(np.int64), 135)
    x = u * 32.0 - (2.0 * k + 1.0)
    c = tab.reshape(136, 8)[k]
    got = np.zeros_like(u)
    for j in range(7, -1, -1):
        got = got * x + c[:, j]
    eps = 1e-12
    p1, p0 = ndtr(u), ndtr(-u)
    want = -(p1 * np.log(p1 + eps) + p0 * np.log(p0 + eps))
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-16)


def test_normal_cdf_table_matches_erfc():
    """phi_tab (csrc/ital_kernels.cuh): (Phi, phi) on the grid x_k = -8.5 + k/128 plus a 5th-order Taylor step from the
    nearest grid point, evaluated as the kernels do, against scipy's ndtr."""
    from scipy.special import ndtr
    lib = _capi.load()
    n = int(lib.ital_phi_table(None, 0))
    tab = np.zeros(n)
    assert lib.ital_phi_table(_capi.dptr(tab), n) == n == 2177 * 2
    rng = np.random.default_rng(6)
    x = np.concatenate((rng.uniform(-8.5, 8.5, 50000), -8.5 + (np.arange(2177) + 0.5) / 128.0 - 1e-12, [-8.5, 8.5, 0.0]))
    x = np.clip(x, -8.5, 8.5)
    k = np.rint((x + 8.5) * 128).astype(np.int64)
    xk = k / 128.0 - 8.5
    dl = x - xk
    t = tab.reshape(2177, 2)[k]
    x2 = xk * xk
    c2, c3, c4, c5 = -0.5 * xk, (x2 - 1) / 6, -xk * (x2 - 3) / 24, (x2 * (x2 - 6) + 3) / 120
    poly = 1 + dl * (c2 + dl * (c3 + dl * (c4 + dl * c5)))
    got = t[:, 0] + t[:, 1] * dl * poly
    np.testing.assert_allclose(got, ndtr(x), rtol=0, atol=3e-16)


@pytest.mark.parametrize('tB,u', [(0, 3), (2, 2), (1, 4), (3, 5), (2, 0)])
def test_subset_node_sets_match_the_oracle(tB, u):
    """csrc/snq_host.h generate_sub (change_estimation_subset) against oracle/ce_subset.py sub_sets: node sets of the
    batch alone and of the prior inside (r_B, s*), conditional moments of the subset given the labels of the batch."""
    from oracle import ce_subset
    lib = _capi.load()
    rng = np.random.RandomState(10 * tB + u)
    D = tB + u
    A = rng.randn(D, D + 2)
    C = A @ A.T / (D + 2) + 0.05 * np.eye(D)
    L = np.ascontiguousarray(np.linalg.cholesky(C))
    m = rng.randn(D) * 0.7
    noise = 1e-4
    sizes = np.zeros(4, dtype=np.int64)
    i32 = ctypes.POINTER(ctypes.c_int32)
    assert lib.ital_snq_sub(tB, D, _capi.dptr(m), _capi.dptr(L), noise, _capi.i64ptr(sizes), None, None, None, None) == 0
    N, NG, bits, nt = [int(x) for x in sizes]
    G = 1 << tB
    assert NG == 2 * G
    eta, w, gb, tab = np.zeros((D, N)), np.zeros(N), np.zeros(NG + 1, dtype=np.int32), np.zeros(max(nt, 1))
    assert lib.ital_snq_sub(tB, D, _capi.dptr(m), _capi.dptr(L), noise, _capi.i64ptr(sizes), _capi.dptr(eta),
                            _capi.dptr(w), gb.ctypes.data_as(i32), _capi.dptr(tab)) == 0
    S = ce_subset.sub_sets(tB, m, L, noise)
    assert bits == S['sub_bits']
    for g in range(G):
        for part, key in ((0, 'part1'), (1, 'part2')):
            e, ww = S[key][g]
            lo, hi = gb[part * G + g], gb[part * G + g + 1]
            assert hi - lo == len(ww)
            np.testing.assert_allclose(w[lo:hi], ww, rtol=1e-9, atol=1e-300)
            np.testing.assert_allclose(eta[:, lo:hi].T, e, rtol=0, atol=1e-9)
    off = 0
    for key, count in (('mass1', G), ('mass2', G), ('mu', G * D), ('Sig', D * D), ('mU', G * u), ('CU', u * u),
                       ('BS', u * D)):
        np.testing.assert_allclose(tab[off:off + count], np.asarray(S[key]).reshape(-1), rtol=1e-9, atol=1e-12,
                                   err_msg=key)
        off += count
    assert off == nt


def test_subset_oracle_forms_agree():
    """oracle/ce_subset.py: the shared-node form (what k_eval_sub computes) against the literal restatement of
    MutualInformation._call_iter_sub (three orthant probabilities per relevance configuration, updated_prediction)."""
    from oracle.ce_subset import mi_sub_literal
    from oracle.ital_oracle import OracleITAL
    rng = np.random.RandomState(0)
    X = rng.randn(40, 2)
    L = OracleITAL(X, length_scale=1.0, noise=1e-6, change_estimation_subset=3)
    L.update({0: 1, 5: -1, 9: 1})
    np.random.seed(7)
    ret = L.fetch_unlabelled(2)
    B = []
    for it, tr in enumerate(L.trace):
        ext = L.subset + [b for b in B if b not in L.subset]
        members = [int(np.nonzero(tr['candidates'] == s)[0][0]) for s in L.subset if s in tr['candidates']]
        for pos in list(range(0, len(tr['candidates']), 9)) + members:
            i = int(tr['candidates'][pos])
            if i in ext:
                r, rel_it = ext, [ext.index(b) for b in B] + [ext.index(i)]
            else:
                r, rel_it = ext + [i], [ext.index(b) for b in B] + [len(ext)]
            lit = mi_sub_literal(L, r, rel_it)
            assert abs(lit - tr['scores'][pos]) <= 2e-5 * max(1.0, abs(lit)), (it, i, lit, tr['scores'][pos])
        B.append(tr['chosen'])
    assert ret == B
