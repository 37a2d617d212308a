"""General feedback model (label_prob < 1 and / or mistake_prob > 0) with SHARED conditional node sets -- ORACLE
(test infrastructure, not product code).

``OracleITAL._mi_general`` restates MutualInformation._call_iter_all literally (/root/reference/ital/ital.py:183-224):
for every relevance configuration r and every feedback configuration f it conditions the block's posterior on the
annotated samples (updated_prob_rel, ital.py:432-450 -> gp.updated_prediction, gp.py:295-344) and evaluates one
orthant probability with its own quadrature nodes.  This module evaluates the SAME sum with node sets that depend only
on the base (the samples already in the batch), so that one set serves every candidate -- the form the CUDA kernel
k_eval_general uses (ital_b200/csrc/snq_host.h generate_general).  With D = t + 1 samples, O the annotated subset,

    MI = sum_r p_r { sum_{O != 0} (1-lp)^(D-|O|) lp^|O| [ (1-mp)^|O| log(q_{r,O} + eps) + (1 - (1-mp)^|O|) log eps ]
                     - (1 - (1-lp)^D) log(p_r + eps) },

where q_{r,O} = P(unannotated samples keep the signs of r | annotated samples labelled as in r) and a label that
contradicts r leaves probability ~0 for r (variances >> label noise).  For every annotated subset O_b of the base and
every sign pattern of its labels, the whitened base coordinates eta are Gaussian N(mu, Sigma) with
mu = A^T (A A^T + noise I)^-1 (f - m_O), Sigma = I - A^T (A A^T + noise I)^-1 A, A = L[O_b, :]; the unannotated base
variables get shared-node quadrature nodes of their own (oracle/orthant.py), mapped back to eta.  An unannotated
candidate contributes sum w Phi(.) per (set, orthant); an annotated candidate contributes the Gaussian density of
its label, normalised per set (Bayes).

tests/test_oracle_golden.py checks this form against the literal enumeration (same quantity, different node
placement: a few 1e-6 relative) and against the goldens recorded from the unmodified reference.
"""
import itertools

import numpy as np
from scipy.special import ndtr

from .orthant import safe_cholesky, snq_nodes, snq_order

EPS = 1e-12


def conditional_sets(m_b, L, noise):
    """Node sets for every annotated subset O_b of the base and every sign pattern of its feedback.

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


def mi_general_shared(m_b, L, m_c, l_c, s_c, label_prob, mistake_prob, noise, estimation='mean'):
    """MI of base + candidate for many candidates: m_c (n,), l_c (n, t), s_c (n,) -> (n,).

    `estimation`: 'mean' (ital.py:216-219), or 'optimistic' / 'pessimistic' (ital.py:210-215: the largest / the smallest
    single term log(q + eps) - log(p_r + eps), weighted by the likelihood of the feedback only, in the enumeration
    order of itertools.product).
    """
    t = len(m_b)
    D = t + 1
    n = len(m_c)
    lp, mp = label_prob, mistake_prob
    sets = conditional_sets(np.asarray(m_b, dtype=np.float64), np.asarray(L, dtype=np.float64).reshape(t, t), noise)
    key = {(s['O'], s['f']): s for s in sets}
    st = np.sqrt(s_c ** 2 + noise)
    tabs = {}
    for (O, f), s in key.items():
        arg = m_c[:, None] + l_c @ s['eta'].T                  # (n, N)
        ng = 1 << len(s['U'])
        onehot = (s['grp'][None, :] == np.arange(ng)[:, None]).astype(np.float64) * s['w'][None, :]   # (ng, N)
        with np.errstate(divide='ignore', invalid='ignore'):
            z = np.where(s_c[:, None] > 0, arg / np.where(s_c > 0, s_c, 1.0)[:, None], np.where(arg > 0, np.inf, -np.inf))
        A = ndtr(z) @ onehot.T                                 # candidate positive, per group
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
    if estimation != 'mean':
        raise NotImplementedError('the shared-node form is restated for label_estimation="mean" only')
    return mi
