"""change_estimation_subset (ITAL(change_estimation_subset = c > 0)) -- ORACLE (test infrastructure, not product code).

The reference (ital/ital.py:102-108, AppendedMutualInformation.__call__ ital.py:514-527 and set_ret/append
ital.py:538-584, MutualInformation._call_iter_sub ital.py:227-275) estimates the change of the model output on a random
subset S of the unseen samples: the relevance of the batch and of the candidate is integrated over, the subset stays at
the signs s* of its means,

    MI(i) = sum_{r over batch + candidate} p_r [ log(P(s*, r | feedback as in r) + eps) - log(P(s*, r) + eps) ],

p_r the probability of r with the subset marginalised (ital.py:259-273; users who label everything without mistakes:
one feedback configuration of likelihood 1 per r, ital.py:313-315).

Two restatements:

* ``mi_sub_literal`` follows the reference call by call for ONE candidate: three orthant probabilities per r, each
  with quadrature nodes of its own (``orthant_prob_any``), the conditioning by ``updated_prediction``.
* ``mi_sub_shared`` evaluates the same sum for many candidates in the form of the CUDA kernel k_eval_sub: p_r and
  P(s*, r) with node sets that depend on ext = batch + subset only (``sub_sets`` restates csrc/snq_host.h
  generate_sub); as in the general feedback model (oracle/general_sets.py) a labelled sample is taken to keep the sign
  of its label (variances >> label noise), so P(s*, r | feedback) = P(S' keeps s* | labels of the batch and the
  candidate), the orthant probability of S' after a rank-one update by the candidate's label, evaluated per
  candidate with the shared-node rule (``orthant_prob``).
"""
import itertools

import numpy as np
from scipy.special import ndtr

from .orthant import SNQ_SC_FROM, orthant_prob, safe_cholesky, sc_orthant, snq_nodes, snq_order

EPS = 1e-12
SUB_N = 16384          # lattice nodes per orthant of 6 or more variables (csrc/snq_host.h kSubN)


def orthant_nodes(m, L, b, n_lattice=SUB_N):
    """Nodes of N(m, L L^T) inside orthant b, whitened coordinates: up to 5 variables cut out of the tensor rule,
    more than that a lattice (csrc/snq_host.h orthant_nodes)."""
    u = len(m)
    if u == 0:
        return np.zeros((1, 0)), np.ones(1)
    if u < SNQ_SC_FROM:
        eta, w, orth = snq_nodes(m, L, snq_order(u))
        sel = orth == b
        return eta[sel], w[sel]
    return sc_orthant(m, L, b, n_lattice)


def sub_sets(tB, m, L, noise):
    """Node sets of ext = [B (tB variables), S'] with means m (D,) and Cholesky factor L (D, D); see generate_sub.

    Returns dict(sub_bits, part1 = [(eta (N, D), w)] * G, mass1 (G,), part2, mass2, mu (G, D), Sig (D, D), mU (G, u),
    CU (u, u), BS (u, D))."""
    m = np.asarray(m, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    D = len(m)
    u, G = D - tB, 1 << tB
    sub_bits = int(sum(1 << a for a in range(u) if m[tB + a] > 0.0))
    part1, mass1 = [], np.zeros(G)
    if tB == 0:
        part1.append((np.zeros((1, D)), np.ones(1)))
        mass1[0] = 1.0
    else:
        eta, w, orth = snq_nodes(m[:tB], L[:tB, :tB], snq_order(tB) or None)
        for g in range(G):
            sel = orth == g
            e = np.zeros((int(sel.sum()), D))
            e[:, :tB] = eta[sel]
            part1.append((e, w[sel]))
            mass1[g] = w[sel].sum()
    A = L[:tB, :]
    if tB > 0:
        X = np.linalg.solve(A @ A.T + noise * np.eye(tB), A)          # (tB, D)
    else:
        X = np.zeros((0, D))
    Sig = np.eye(D) - A.T @ X
    Bm = L[tB:, :]
    CU = Bm @ Sig @ Bm.T
    BS = Bm @ Sig
    part2, mass2, mu_all, mU_all = [], np.zeros(G), np.zeros((G, D)), np.zeros((G, u))
    prior = snq_nodes(m, L, snq_order(D)) if D < SNQ_SC_FROM else None
    for g in range(G):
        b = g | (sub_bits << tB)
        if prior is not None:
            sel = prior[2] == b
            e, w = prior[0][sel], prior[1][sel]
        else:
            e, w = orthant_nodes(m, L, b)
        part2.append((e, w))
        mass2[g] = w.sum()
        f = np.array([1.0 if (g >> a) & 1 else -1.0 for a in range(tB)])
        mu = X.T @ (f - m[:tB])
        mu_all[g] = mu
        mU_all[g] = m[tB:] + Bm @ mu
    return dict(sub_bits=sub_bits, part1=part1, mass1=mass1, part2=part2, mass2=mass2, mu=mu_all, Sig=Sig,
                mU=mU_all, CU=CU, BS=BS)


def mi_sub_shared(tB, m_ext, L_ext, m_c, l_c, v_c, noise, mistake_prob=0.0):
    """Scores of many candidates: m_c (n,), l_c (n, D) projections on L_ext, v_c (n,) posterior variances -> (n,)."""
    m_c = np.asarray(m_c, dtype=np.float64)
    l_c = np.asarray(l_c, dtype=np.float64).reshape(len(m_c), -1)
    D = l_c.shape[1]
    G = 1 << tB
    S = sub_sets(tB, m_ext, L_ext, noise)
    s2B = v_c - (l_c[:, :tB] ** 2).sum(axis=1)
    s2F = v_c - (l_c ** 2).sum(axis=1)
    sB, sF = np.sqrt(np.maximum(s2B, 0.0)), np.sqrt(np.maximum(s2F, 0.0))
    st2 = np.maximum(s2F, 0.0) + noise

    def cdf_sum(nodes, sd):
        e, w = nodes
        num = m_c[:, None] + l_c @ e.T
        with np.errstate(divide='ignore', invalid='ignore'):
            z = np.where(sd[:, None] > 0, num / np.where(sd > 0, sd, 1.0)[:, None], np.where(num > 0, np.inf, -np.inf))
        return ndtr(z) @ w

    lSl = np.einsum('nd,de,ne->n', l_c, S['Sig'], l_c)
    tau2 = st2 + np.maximum(lSl, 0.0)
    u = D - tB
    sstar = [(S['sub_bits'] >> a) & 1 for a in range(u)]
    cv = l_c @ S['BS'].T                                   # (n, u): covariance of S' with the candidate's label
    mi = np.zeros(len(m_c))
    c1 = (1.0 - mistake_prob) ** (tB + 1)
    for g in range(G):
        A1 = cdf_sum(S['part1'][g], sB)
        A2 = cdf_sum(S['part2'][g], sF)
        mean_c = m_c + l_c @ S['mu'][g]
        for rc in (0, 1):
            y = 1.0 if rc else -1.0
            p_r = np.maximum(A1 if rc else S['mass1'][g] - A1, 0.0)
            P = np.maximum(A2 if rc else S['mass2'][g] - A2, 0.0)
            if u == 0:
                q = np.ones(len(m_c))
            else:
                q = np.empty(len(m_c))
                for c in range(len(m_c)):                  # rank-one update by the candidate's label
                    mean_u = S['mU'][g] + cv[c] * (y - mean_c[c]) / tau2[c]
                    cov_u = S['CU'] - np.outer(cv[c], cv[c]) / tau2[c]
                    q[c] = orthant_prob(sstar, mean_u, cov_u, snq_order(u - 1) if u > 1 else None)
            q = np.clip(q, 0.0, 1.0)
            # a user who mislabels (fb_iter / likelihood, ital.py:317-328, 453-481): any wrong label contradicts r
            mi += p_r * (c1 * np.log(q + EPS) + (1.0 - c1) * np.log(EPS) - np.log(P + EPS))
    return mi


def orthant_prob_any(rel, mean, cov, n_lattice=65536):
    """P(sign(z) = rel), z ~ N(mean, cov), any number of variables: the last one analytic, the others by the tensor
    rule (up to 5) or a lattice inside their orthant -- the quantity of MutualInformation.prob_rel (ital.py:345-383)."""
    rel = np.asarray(rel).astype(bool)
    mean = np.asarray(mean, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64).reshape(len(mean), len(mean))
    D = len(mean)
    t = D - 1
    sgn = 1.0 if rel[t] else -1.0
    if t == 0:
        sd = np.sqrt(max(cov[0, 0], 0.0))
        return float(ndtr(sgn * mean[0] / sd)) if sd > 0 else float((mean[0] > 0) == rel[0])
    L = safe_cholesky(cov[:t, :t])
    l = np.linalg.solve(L, cov[:t, t])
    s2 = cov[t, t] - l @ l
    b = int(sum(int(r) << j for j, r in enumerate(rel[:t])))
    eta, w = orthant_nodes(mean[:t], L, b, n_lattice)
    num = mean[t] + eta @ l
    if s2 > 0:
        return float(w @ ndtr(sgn * num / np.sqrt(s2)))
    return float(w @ ((num > 0) == rel[t]))


def mi_sub_literal(learner, ret, rel_it, n_lattice=65536, mistake_prob=0.0):
    """MutualInformation._call_iter_sub (ital.py:227-275) for one candidate, call by call.  ``learner``: OracleITAL
    (rel_mean, gp.predict_stored, updated_prediction); ret: sample indices; rel_it: positions in ret to integrate over.
    mistake_prob > 0: the feedback configurations of a user who labels everything but errs (fb_iter ital.py:317-328,
    likelihood ital.py:453-481)."""
    ret = [int(i) for i in ret]
    mean = learner.rel_mean[ret]
    cov = learner.gp.predict_stored(ret, cov_mode='full')[1]
    mean_it, cov_it = mean[rel_it], cov[np.ix_(rel_it, rel_it)]
    rel_vec = mean > 0
    order = sorted(range(len(ret)), key=lambda a: ret[a])                       # updated_prob_rel sorts by index
    mi = 0.0
    for reli in itertools.product([False, True], repeat=len(rel_it)):
        rv = rel_vec.copy()
        rv[rel_it] = reli
        pr = orthant_prob_any(reli, mean_it, cov_it, n_lattice)
        log_pr = np.log(orthant_prob_any(rv, mean, cov, n_lattice) + EPS)
        if mistake_prob > 0:
            fbs = list(itertools.product([-1, 1], repeat=len(rel_it)))
        else:
            fbs = [tuple(1 if r else -1 for r in reli)]
        for fbi in fbs:
            like = 1.0
            if mistake_prob > 0:
                for r, f in zip(reli, fbi):
                    like *= (1.0 - mistake_prob) if (f > 0) == bool(r) else mistake_prob
            feedback = {ret[i]: f for i, f in zip(rel_it, fbi)}
            mean_u, cov_u = learner.updated_prediction(feedback, [ret[a] for a in order])
            pr_updated = orthant_prob_any(rv[order], mean_u, cov_u, n_lattice)
            mi += pr * like * (np.log(pr_updated + EPS) - log_pr)
    return mi
