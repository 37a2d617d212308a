"""numpy prototype of the shared-node evaluation of the GENERAL feedback model (label_prob < 1, any mistake_prob)
used to check the derivation in DESIGN.md against the oracle's literal double loop before it was written in CUDA
(ital_b200/csrc: k_eval_general + snq_host.h conditional node sets).  Development aid only; not product code.

    python tools/proto_general.py
"""
import itertools
import os
import sys

import numpy as np
from scipy.special import ndtr

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.orthant import snq_nodes, snq_order, safe_cholesky  # noqa: E402

EPS = 1e-12


def conditional_sets(m_b, L, noise):
    """Node sets for every observed subset O_b of the base and every sign pattern of its feedback.

    Returns a list of dicts: O (tuple of base positions), f (tuple of +-1), U (tuple), eta (N, t), w (N,),
    grp (N,) orthant id over U (bit k = variable U[k] positive).
    """
    t = len(m_b)
    sets = []
    for k in range(0, t + 1):
        for O in itertools.combinations(range(t), k):
            U = tuple(j for j in range(t) if j not in O)
            for f in itertools.product((-1.0, 1.0), repeat=k):
                if k == 0:
                    mu, Sig = np.zeros(t), np.eye(t)
                else:
                    A = L[list(O), :]
                    S = A @ A.T + noise * np.eye(k)
                    Ki = np.linalg.solve(S, A)                   # (k, t)
                    mu = Ki.T @ (np.array(f) - m_b[list(O)])
                    Sig = np.eye(t) - A.T @ Ki
                if len(U) == 0:
                    sets.append(dict(O=O, f=f, U=U, eta=mu[None, :], w=np.ones(1), grp=np.zeros(1, dtype=np.int64)))
                    continue
                B = L[list(U), :]
                mU = m_b[list(U)] + B @ mu
                CU = B @ Sig @ B.T
                Lc = safe_cholesky(CU)
                zeta, w, grp = snq_nodes(mU, Lc, snq_order(len(U)))
                G = Sig @ B.T @ np.linalg.inv(Lc).T              # eta = mu + G zeta
                sets.append(dict(O=O, f=f, U=U, eta=mu[None, :] + zeta @ G.T, w=w, grp=grp))
    return sets


def mi_general_shared(m_b, L, m_c, l_c, s_c, label_prob, mistake_prob, noise):
    """MI of base + candidate for many candidates (m_c (n,), l_c (n, t), s_c (n,))."""
    t = len(m_b)
    D = t + 1
    n = len(m_c)
    lp, mp = label_prob, mistake_prob
    sets = conditional_sets(m_b, L, noise)
    key = {(s['O'], s['f']): s for s in sets}
    st = np.sqrt(s_c ** 2 + noise)
    tabs = {}
    for (O, f), s in key.items():
        arg = m_c[:, None] + l_c @ s['eta'].T                  # (n, N)
        ng = 1 << len(s['U'])
        onehot = (s['grp'][None, :] == np.arange(ng)[:, None]).astype(np.float64) * s['w'][None, :]   # (ng, N)
        A = ndtr(arg / s_c[:, None]) @ onehot.T                # candidate positive, per group
        Wg = onehot.sum(axis=1)
        Bp = np.exp(-0.5 * ((1.0 - arg) / st[:, None]) ** 2) @ onehot.T
        Bm = np.exp(-0.5 * ((-1.0 - arg) / st[:, None]) ** 2) @ onehot.T
        tabs[(O, f)] = (A, Wg, Bp, Bm)
    A0, W0, _, _ = tabs[((), ())]
    mi = np.zeros(n)
    for r in itertools.product((0, 1), repeat=D):
        rb, rc = r[:t], r[t]
        g0 = sum(rb[j] << j for j in range(t))
        p_r = A0[:, g0] if rc else W0[g0] - A0[:, g0]
        p_r = np.maximum(p_r, 0.0)
        inner = -(1.0 - (1.0 - lp) ** D) * np.log(p_r + EPS)
        for k in range(1, D + 1):
            for Ofull in itertools.combinations(range(D), k):
                lam = (1.0 - lp) ** (D - k) * lp ** k
                c_in = t in Ofull
                Ob = tuple(j for j in Ofull if j < t)
                fb = tuple(2.0 * rb[j] - 1.0 for j in Ob)
                A, Wg, Bp, Bm = tabs[(Ob, fb)]
                Ub = tuple(j for j in range(t) if j not in Ob)
                g = sum(rb[j] << kk for kk, j in enumerate(Ub))
                if k == D:
                    q = np.ones(n)
                elif not c_in:
                    q = A[:, g] if rc else Wg[g] - A[:, g]
                else:
                    Bsel = Bp if rc else Bm
                    q = Bsel[:, g] / np.maximum(Bsel.sum(axis=1), 1e-300)
                q = np.clip(q, 0.0, 1.0)
                inner = inner + lam * ((1.0 - mp) ** k * np.log(q + EPS) + (1.0 - (1.0 - mp) ** k) * np.log(EPS))
        mi += p_r * inner
    return mi


def main():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
    from conftest import drive, load_golden
    from oracle.ital_oracle import OracleITAL
    import scipy.linalg
    for name in ('toy_mistakes_k3', 'butterflies_conservative_k3'):
        g = load_golden(name)
        kw = dict(g['learner_kw'])
        ora = drive(OracleITAL(g['X'], **kw), g)
        ora.fetch_unlabelled(int(g['k']), forced=g['ret'].tolist())
        for t, (tr, st) in enumerate(zip(ora.trace, g['steps'])):
            m_b = ora.rel_mean[g['ret'][:t]] if t else np.zeros(0)
            if t:
                L = safe_cholesky(tr['cov_base'])
                l = scipy.linalg.solve_triangular(L, tr['cov_base_test'], lower=True).T
            else:
                L = np.zeros((0, 0))
                l = np.zeros((len(tr['candidates']), 0))
            s = np.sqrt(np.maximum(tr['var'] - (l * l).sum(axis=1), 0.0))
            mi = mi_general_shared(m_b, L, tr['mean'], l, s, kw['label_prob'], kw['mistake_prob'], kw['noise'])
            err_o = np.abs(mi - tr['scores']) / np.abs(tr['scores'])
            err_g = np.abs(mi - st['mi']) / np.abs(st['mi'])
            print('%-28s step %d: max rel err vs oracle %.2e, vs reference golden %.2e, argmax %s/%s/%s'
                  % (name, t, err_o.max(), err_g.max(), tr['candidates'][np.argmax(mi)],
                     tr['candidates'][np.argmax(tr['scores'])], st['chosen']))


if __name__ == '__main__':
    main()
