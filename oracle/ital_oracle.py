"""CPU float64 restatement of ITAL's batch-selection hot path -- ORACLE (test infrastructure only).

This file restates, in numpy, what the reference computes on the path
``ITAL.fetch_unlabelled`` -> ``AppendedMutualInformation`` -> ``MutualInformation`` -> ``GaussianProcess``
(/root/reference/ital/ital.py:84-134, 183-224, 278-383, 432-481, 485-586; /root/reference/ital/gp.py:8-87,
141-261, 295-344, 390-416; /root/reference/ital/retrieval_base.py:34-194).  It is NOT the product: only
tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs import it,
as the checker.  The product path (ital_b200/) never touches it and has no CPU fallback.

Differences from the reference that are deliberate and documented in DESIGN.md:

* The reference materialises the n-by-n kernel matrix ``K_all`` (gp.py:128).  Every ``K_all[np.ix_(..)]``
  gather is restated here as a Gram of data rows against the few rows involved (columns on demand), which is
  the same arithmetic (gp.py:414-416) evaluated only where it is used.
* Orthant probabilities of two or more variables come from ``oracle/orthant.py`` (fixed-node rule) instead
  of the randomised third-party ``MVNDST``.  Pinning: one variable is the reference's closed form; two
  variables agree with Genz's BVU to ~1e-12; three and more are checked against the reference code driven
  through an accurate ``mvndst`` stand-in (tests/golden/make_golden.py).  Against the historical Fortran
  routine (1e-4 noise) parity for three or more variables is **unpinned**.
* Monte-Carlo modes, ``clip_cov`` grouping and ``change_estimation_subset`` (ital.py:227-275, 293-297,
  386-429) are not restated: none of the named configurations uses them.
"""

import itertools

import numpy as np
import scipy.linalg.lapack

from .orthant import orthant_prob_all, safe_cholesky, snq_joint, snq_order

EPS = 1e-12   # MutualInformation.__init__ default (ital.py:144)


def invh(M):
    """Cholesky inverse via dpotrf/dpotri, symmetrised (gp.py:8-37)."""
    zz, _ = scipy.linalg.lapack.dpotrf(M, False, False)
    inv_M, _ = scipy.linalg.lapack.dpotri(zz)
    i, j = np.triu_indices_from(inv_M, k=1)
    inv_M[j, i] = inv_M[i, j]
    return inv_M


class OracleGP(object):
    """GaussianProcess (gp.py:91-436) without the n-by-n matrix."""

    def __init__(self, data, length_scale, var=1.0, noise=1e-6):
        self.X = np.array(data, dtype=np.float64)
        self.length_scale = length_scale
        self.length_scale_sq = length_scale * length_scale
        self.var = var
        self.noise = noise
        self.sqnorm = np.sum(self.X ** 2, axis=-1)          # gp.py:411
        self.reset()

    def reset(self):                                           # gp.py:132-138
        self.ind = []
        self.y = self.K = self.K_inv = self.w = None
        self._cols = {}

    def kernel(self, a, b):                                    # gp.py:390-416
        a = np.atleast_2d(np.asarray(a, dtype=np.float64))
        b = np.atleast_2d(np.asarray(b, dtype=np.float64))
        a_norm = np.sum(a ** 2, axis=-1)
        b_norm = np.sum(b ** 2, axis=-1)
        s = -2 * self.length_scale_sq
        return self.var * np.exp((a_norm[:, None] + b_norm[None, :] - 2 * np.dot(a, b.T)) / s)

    def col(self, i):
        """K_all[:, i] on demand (one column of gp.py:128)."""
        i = int(i)
        if i not in self._cols:
            s = -2 * self.length_scale_sq
            self._cols[i] = self.var * np.exp((self.sqnorm + self.sqnorm[i] - 2 * (self.X @ self.X[i])) / s)
        return self._cols[i]

    def cols(self, ind):
        """K_all[:, ind] as an N-by-len(ind) array."""
        if len(ind) == 0:
            return np.zeros((len(self.X), 0))
        return np.stack([self.col(i) for i in ind], axis=1)

    def block(self, a, b):
        """K_all[np.ix_(a, b)]."""
        return self.cols(b)[np.asarray(a, dtype=np.int64)] if len(a) else np.zeros((0, len(b)))

    def fit(self, ind, y):                                     # gp.py:141-161
        self.ind = [int(i) for i in ind]
        self.y = np.array(y, dtype=np.float64)
        self.K = self.block(self.ind, self.ind) + self.noise * np.eye(len(self.ind))
        self.K_inv = invh(self.K)
        self.w = np.dot(self.K_inv, self.y)
        return self

    def update(self, ind, y):                                  # gp.py:164-200
        if len(self.ind) == 0:
            return self.fit(ind, y)
        ind = [int(i) for i in ind]
        y = np.asarray(y, dtype=np.float64)
        K_old_new = self.block(self.ind, ind)
        K_new = self.block(ind, ind) + self.noise * np.eye(len(ind))
        self.ind += ind
        self.y = np.concatenate((self.y, y))
        self.K = np.vstack((np.hstack((self.K, K_old_new)), np.hstack((K_old_new.T, K_new))))
        self.K_inv = invh(self.K)
        self.w = np.dot(self.K_inv, self.y)
        return self

    def predict_stored(self, ind=None, cov_mode=None):         # gp.py:203-232
        k_test = self.cols(self.ind).T if ind is None else self.block(self.ind, ind)
        pred_mean = np.dot(self.w.T, k_test)
        if cov_mode == 'full':
            if ind is None:
                raise ValueError('full covariance over all samples is the n-by-n matrix this oracle avoids')
            return pred_mean, self.block(ind, ind) - np.dot(k_test.T, np.dot(self.K_inv, k_test))
        elif cov_mode == 'diag':
            return pred_mean, np.maximum(0, self.var - np.sum(k_test * np.dot(self.K_inv, k_test), axis=0))
        return pred_mean

    def updated_prediction(self, ind, y, pred_ind, cov_mode=None):     # gp.py:295-344
        """Prediction for rows `pred_ind` after hypothetically adding (ind, y); the model itself is unchanged.
        The reference extends K^-1 by a Woodbury identity (extend_inv, gp.py:40-87); the same extended inverse is
        formed here from the Schur complement of the new block."""
        ind = [int(i) for i in ind]
        pred_ind = [int(i) for i in pred_ind]
        B = self.block(self.ind, ind)                                   # old x new
        D = self.block(ind, ind) + self.noise * np.eye(len(ind))
        KiB = np.dot(self.K_inv, B)
        S_inv = np.linalg.inv(D - np.dot(B.T, KiB))
        K_inv = np.vstack((np.hstack((self.K_inv + KiB @ S_inv @ KiB.T, -KiB @ S_inv)),
                           np.hstack((-S_inv @ KiB.T, S_inv))))
        k_test = self.block(self.ind + ind, pred_ind)
        pred_mean = np.dot(np.dot(K_inv, np.concatenate((self.y, np.asarray(y, dtype=np.float64)))).T, k_test)
        if cov_mode == 'full':
            return pred_mean, self.block(pred_ind, pred_ind) - np.dot(k_test.T, np.dot(K_inv, k_test))
        elif cov_mode == 'diag':
            return pred_mean, np.maximum(0, self.var - np.sum(k_test * np.dot(K_inv, k_test), axis=0))
        return pred_mean

    def predict_cov_parts(self, base_ind):
        """The three ingredients of predict_cov_batch (gp.py:250-256) for ind = all rows."""
        base_ind = [int(i) for i in base_ind]
        k_base = self.block(self.ind, base_ind)
        cov_base = self.block(base_ind, base_ind) - np.dot(k_base.T, np.dot(self.K_inv, k_base))
        k_test = self.cols(self.ind).T
        var_test = self.var - np.sum(k_test * np.dot(self.K_inv, k_test), axis=0)
        cov_base_test = self.cols(base_ind).T - np.dot(k_base.T, np.dot(self.K_inv, k_test))
        return cov_base, var_test, cov_base_test

    def predict(self, X, cov_mode=None):                       # gp.py:264-292
        k_test = self.kernel(self.X[self.ind], X)
        pred_mean = np.dot(self.w.T, k_test)
        if cov_mode == 'full':
            return pred_mean, self.kernel(X, X) - np.dot(k_test.T, np.dot(self.K_inv, k_test))
        elif cov_mode == 'diag':
            return pred_mean, np.maximum(0, self.var - np.sum(k_test * np.dot(self.K_inv, k_test), axis=0))
        return pred_mean


def entropy_terms(p, p_upd=1.0, eps=EPS):
    """p * (log(p' + eps) - log(p + eps)), the summand of ital.py:207-219 ('mean' estimation)."""
    return p * (np.log(p_upd + eps) - np.log(p + eps))


def mi_perfect_user(m_base, cov_base, m_c, var_c, cov_base_c, q=None, pool=None):
    """MI of ``ret + [i]`` for every candidate i under the perfect-user model, vectorised.

    With label_prob >= 1 and mistake_prob <= 0 there is one feedback configuration per relevance
    configuration (ital.py:313-315) and the updated orthant probability is 1 to double precision whenever
    the posterior variances are large against the label noise, so
    ``MI = sum_r p_r * (log(1 + eps) - log(p_r + eps))``.  Returns (scores, p_plus, p_base,
    flagged) where ``flagged`` marks candidates whose conditional variance is too small for that shortcut.
    """
    t = len(m_base)
    m_c = np.asarray(m_c, dtype=np.float64)
    if t == 0:
        sd = np.sqrt(np.maximum(var_c, 0.0))
        with np.errstate(divide='ignore', invalid='ignore'):
            z = np.where(sd > 0, m_c / np.where(sd > 0, sd, 1.0), np.where(m_c > 0, np.inf, -np.inf))
        from scipy.special import ndtr
        p1 = ndtr(z)
        p0 = ndtr(-z)
        return entropy_terms(p0) + entropy_terms(p1), p1[:, None], np.ones(1), sd
    L = safe_cholesky(cov_base)
    l = scipy.linalg.solve_triangular(L, cov_base_c, lower=True).T        # (N, t)
    s2 = var_c - np.sum(l * l, axis=1)
    s = np.sqrt(np.maximum(s2, 0.0))
    p_plus, p_base = snq_joint(m_base, L, m_c, l, s, q, pool=pool)
    p_minus = np.maximum(p_base[None, :] - p_plus, 0.0)
    scores = entropy_terms(p_plus).sum(axis=1) + entropy_terms(p_minus).sum(axis=1)
    return scores, p_plus, p_base, s


def _joint_cov(cov_base, var_c, cov_base_c):
    D = len(cov_base) + 1
    cov = np.empty((D, D))
    cov[:D - 1, :D - 1] = cov_base
    cov[:D - 1, D - 1] = cov[D - 1, :D - 1] = cov_base_c
    cov[D - 1, D - 1] = var_c
    return cov


def group_cov(corr, th):
    """Connected components of the graph |corr| > th (group_cov, ital.py:590-616), in the reference's order: groups by
    their smallest unassigned member, members in order of discovery."""
    clipped = np.abs(corr) > th
    unassigned = list(range(len(corr)))
    groups = []
    while unassigned:
        group, frontier = [], [j for j in np.nonzero(clipped[unassigned[0]])[0]]
        while frontier:
            group += [j for j in frontier if j not in group]
            unassigned = [j for j in unassigned if j not in frontier]
            reach = np.nonzero(clipped[group].max(axis=0))[0]
            frontier = [j for j in reach if j in unassigned]
        groups.append(group)
    return groups


def mi_grouped(mean, cov, th, eps=EPS):
    """MI of more than 5 samples with clip_cov = th under the perfect-user model (prob_rel -> _grouped_prob_rel,
    ital.py:360-429): correlations below th are dropped, the probability of a relevance configuration is the product
    over the independent groups, and the sum over all 2^D configurations is taken literally (ital.py:193-219)."""
    sd = np.sqrt(np.diag(cov))
    corr = cov / np.outer(sd, sd)
    groups = group_cov(corr, th)
    D = len(mean)
    tables = []
    groups = [sorted(g) for g in groups]        # (batch order, the candidate last: the variable integrated analytically)
    for g in groups:
        if len(g) == 1:
            j = g[0]
            p0 = float(ndtr_scalar(-mean[j] / sd[j]))
            tables.append(np.array([p0, 1.0 - p0]))
        else:
            tables.append(orthant_prob_all(mean[g], cov[np.ix_(g, g)], snq_order(len(g) - 1) or None))
    p = np.ones(1 << D)
    r = np.arange(1 << D)
    for g, tab in zip(groups, tables):
        idx = np.zeros(1 << D, dtype=np.int64)
        for k, j in enumerate(g):
            idx |= ((r >> j) & 1) << k
        p = p * tab[idx]
    return float(np.sum(entropy_terms(p, 1.0, eps)))


def ndtr_scalar(x):
    from scipy.special import ndtr
    return ndtr(x)


class OracleITAL(object):
    """ITAL + ActiveRetrievalBase (ital.py:12-134, retrieval_base.py:7-194) over OracleGP."""

    def __init__(self, data=None, queries=[], length_scale=0.1, var=1.0, noise=1e-6,
                 label_prob=1.0, mistake_prob=0.0, top_candidates=None, change_estimation_subset=0,
                 clip_cov=0, label_estimation='mean', monte_carlo_num_rel=None, monte_carlo_num_fb=None,
                 parallelized=True, force_general=False, general_sets=False):
        self.length_scale, self.var, self.noise = length_scale, var, noise
        self.label_prob, self.mistake_prob = label_prob, mistake_prob
        self.top_candidates = top_candidates
        self.label_estimation = label_estimation
        self.parallelized = parallelized
        self.force_general = force_general
        self.general_sets = general_sets      # general feedback model through shared conditional node sets (oracle/general_sets.py)
        self.change_estimation_subset = change_estimation_subset
        self.clip_cov = clip_cov
        if change_estimation_subset is None or monte_carlo_num_rel is not None or monte_carlo_num_fb is not None:
            raise NotImplementedError('oracle restates the enumeration path only (see module docstring)')
        self.fit(data, queries)

    # ---- retrieval_base.py ---------------------------------------------------------------------------
    def fit(self, data, queries=[]):                                            # retrieval_base.py:34-45
        self.data = data
        self.queries = queries
        if self.data is not None:
            X = np.concatenate((self.data, self.queries)) if len(self.queries) > 0 else self.data
            self.gp = OracleGP(X, self.length_scale, self.var, self.noise)
            self.reset()
        else:
            self.gp = None

    def reset(self):                                                            # retrieval_base.py:48-61
        self.rounds = 0
        self.relevant_ids, self.irrelevant_ids, self.unnameable_ids = set(), set(), set()
        if len(self.queries) > 0:
            n = len(self.data)
            self.gp.fit(np.arange(n, n + len(self.queries)), [1] * len(self.queries))
            self.rel_mean = self.gp.predict_stored()[:n]
        else:
            self.gp.reset()
            self.rel_mean = None

    def top_results(self, k=None):                                              # retrieval_base.py:64-75
        ind = np.argsort(self.rel_mean)[::-1]
        return ind[:k] if k is not None else ind

    def get_unseen(self):                                                       # retrieval_base.py:78-87
        seen = self.relevant_ids | self.irrelevant_ids | self.unnameable_ids
        return [i for i in range(len(self.data)) if i not in seen]

    def partition_feedback(self, feedback):                                     # retrieval_base.py:167-194
        rel, irr, unnameable = [], [], []
        for i, fb in feedback.items():
            if fb > 0:
                if i in self.irrelevant_ids:
                    raise RuntimeError('Cannot change feedback once given.')
                elif i not in self.relevant_ids:
                    rel.append(i)
            elif fb < 0:
                if i in self.relevant_ids:
                    raise RuntimeError('Cannot change feedback once given.')
                elif i not in self.irrelevant_ids:
                    irr.append(i)
            else:
                unnameable.append(i)
        return rel, irr, unnameable

    def update(self, feedback):                                                 # retrieval_base.py:105-126
        rel, irr, unnameable = self.partition_feedback(feedback)
        if len(rel) + len(irr) > 0:
            self.gp.update(rel + irr, np.concatenate((np.ones(len(rel)), -1 * np.ones(len(irr)))))
            self.rel_mean = self.gp.predict_stored()[:len(self.data)]
            self.relevant_ids.update(rel)
            self.irrelevant_ids.update(irr)
            self.rounds += 1
        self.unnameable_ids.update(unnameable)

    # ---- ital.py -------------------------------------------------------------------------------------
    def updated_prediction(self, feedback, test_ind, cov_mode='full'):          # retrieval_base.py:129-164
        rel, irr, _ = self.partition_feedback(feedback)
        if len(rel) + len(irr) == 0:
            return self.gp.predict_stored(test_ind, cov_mode=cov_mode)
        rel.sort()
        irr.sort()
        return self.gp.updated_prediction(rel + irr, np.concatenate((np.ones(len(rel)), -np.ones(len(irr)))),
                                          test_ind, cov_mode=cov_mode)

    def _perfect_user(self):
        return (self.label_prob >= 1) and (self.mistake_prob <= 0)            # ital.py:313

    def fetch_unlabelled(self, k, show_progress=False, forced=None, pool=None):   # ital.py:84-134
        """``forced`` (test hook): follow these choices instead of the argmax so that a recorded greedy path
        can be re-scored step by step even where the maximum is an exact tie.  ``pool``: multiprocessing pool over
        which the candidates of every step are spread (the reference's ``parallelized=True``, ital.py:124-126)."""
        candidates = self.get_unseen()
        if len(candidates) < k:
            k = len(candidates)
        subset = None
        if self.change_estimation_subset:                                       # ital.py:105-106 (same draw, same RNG)
            subset = [int(i) for i in sorted(np.random.choice(
                candidates, min(len(candidates), self.change_estimation_subset), replace=False))]
        if self.top_candidates is not None:                                     # ital.py:111-117
            top = self.top_candidates
            if isinstance(top, float):
                top = min(len(candidates), int(top * (len(self.queries) + len(self.relevant_ids)
                                                      + len(self.irrelevant_ids))))
            if (top > 0) and (top < len(candidates)):
                top_ind = np.argpartition(self.rel_mean[candidates], -top)[-top:]
                candidates = [candidates[i] for i in top_ind]
        if subset is not None:
            return self._fetch_change_subset(k, candidates, subset, forced)
        n = len(self.data)
        ret = []
        self.trace = []          # per greedy step: dict(candidates, scores, ...) for the parity tests
        var0 = self.gp.predict_stored(cov_mode='diag')[1]                       # ital.py:557-558 (clamped)
        for it in range(k):
            cand = np.asarray(candidates, dtype=np.int64)
            if len(ret) == 0:
                cov_base, var_test, cov_base_test = np.zeros((0, 0)), var0, np.zeros((0, len(var0)))
            else:
                cov_base, var_test, cov_base_test = self.gp.predict_cov_parts(ret)   # ital.py:586
            m_base = self.rel_mean[ret] if len(ret) else np.zeros(0)
            if 0 < self.clip_cov < 1 and len(ret) + 1 > 5:                      # ital.py:360-362
                if not (self._perfect_user() and self.label_estimation == 'mean'):
                    raise NotImplementedError('clip_cov is restated for users who label everything correctly')
                scores = np.array([mi_grouped(np.concatenate((m_base, [self.rel_mean[i]])),
                                              _joint_cov(cov_base, var_test[i], cov_base_test[:, i]), self.clip_cov)
                                   for i in cand])
                extra = {}
            elif self._perfect_user() and self.label_estimation == 'mean' and not self.force_general:
                scores, p_plus, p_base, s = mi_perfect_user(
                    m_base, cov_base, self.rel_mean[cand], var_test[cand], cov_base_test[:, cand], pool=pool)
                extra = dict(p_plus=p_plus, p_base=p_base, s=s)
            elif self.general_sets and self.label_estimation == 'mean':
                from .general_sets import mi_general_shared
                L = safe_cholesky(cov_base) if len(ret) else np.zeros((0, 0))
                l = scipy.linalg.solve_triangular(L, cov_base_test[:, cand], lower=True).T if len(ret) \
                    else np.zeros((len(cand), 0))
                s = np.sqrt(np.maximum(var_test[cand] - (l * l).sum(axis=1), 0.0))
                scores = mi_general_shared(m_base, L, self.rel_mean[cand], l, s, self.label_prob, self.mistake_prob,
                                           self.noise)
                extra = dict(s=s)
            else:
                scores = np.array([self._mi_general(ret + [int(i)], m_base, cov_base, self.rel_mean[i],
                                                    var_test[i], cov_base_test[:, i]) for i in cand])
                extra = {}
            max_ind = int(np.argmax(scores))                                    # ital.py:130 (first maximum)
            if forced is not None:
                max_ind = candidates.index(int(forced[it]))
            self.trace.append(dict(candidates=cand, scores=scores, mean=self.rel_mean[cand].copy(),
                                   var=var_test[cand].copy(), cov_base=cov_base.copy(),
                                   cov_base_test=cov_base_test[:, cand].copy(), chosen=int(cand[max_ind]),
                                   argmax=int(cand[int(np.argmax(scores))]), **extra))
            ret.append(int(cand[max_ind]))
            del candidates[max_ind]
        return ret

    # change_estimation_subset > 0 (ital.py:227-275, 514-584), scores in the shared-node form of oracle/ce_subset.py
    def _sub_scores(self, batch, sub, rows):
        """Scores of ``rows`` (outside ext) against ext = batch + sub."""
        from .ce_subset import mi_sub_shared
        ext = list(batch) + list(sub)
        rows = np.asarray(rows, dtype=np.int64)
        var0 = self.gp.predict_stored(cov_mode='diag')[1]
        if len(ext) == 0:
            L, l = np.zeros((0, 0)), np.zeros((len(rows), 0))
        else:
            cov_ext, _, cov_ext_test = self.gp.predict_cov_parts(ext)
            L = safe_cholesky(cov_ext)
            l = scipy.linalg.solve_triangular(L, cov_ext_test[:, rows], lower=True).T
        m_ext = self.rel_mean[ext] if len(ext) else np.zeros(0)
        return mi_sub_shared(len(batch), m_ext, L, self.rel_mean[rows], l, var0[rows], self.noise, self.mistake_prob)

    def _fetch_change_subset(self, k, candidates, subset, forced=None):
        if not (self.label_prob >= 1 and self.label_estimation == 'mean'):
            raise NotImplementedError('change_estimation_subset is restated for users who label everything')
        ret = []
        self.trace = []
        self.subset = list(subset)
        for it in range(k):
            sub = [i for i in subset if i not in ret]
            cand = np.asarray(candidates, dtype=np.int64)
            scores = np.empty(len(cand))
            outside = np.array([int(i) not in sub for i in cand], dtype=bool)
            if outside.any():
                scores[outside] = self._sub_scores(ret, sub, cand[outside])
            for pos in np.nonzero(~outside)[0]:         # a subset member: scored with itself moved out of the subset
                i = int(cand[pos])
                scores[pos] = self._sub_scores(ret, [j for j in sub if j != i], [i])[0]
            max_ind = int(np.argmax(scores))
            if forced is not None:
                max_ind = candidates.index(int(forced[it]))
            self.trace.append(dict(candidates=cand, scores=scores, chosen=int(cand[max_ind]),
                                   argmax=int(cand[int(np.argmax(scores))]), subset=list(subset)))
            ret.append(int(cand[max_ind]))
            del candidates[max_ind]
        return ret

    # General feedback model: literal restatement of _call_iter_all for one candidate (ital.py:183-224).
    def _mi_general(self, ids, m_base, cov_base, m_c, var_c, cov_base_c):
        D = len(ids)
        mean = np.concatenate((m_base, [m_c]))
        cov = np.empty((D, D))
        cov[:D - 1, :D - 1] = cov_base
        cov[:D - 1, D - 1] = cov[D - 1, :D - 1] = cov_base_c
        cov[D - 1, D - 1] = var_c
        p_all = orthant_prob_all(mean, cov, snq_order(D - 1) or None)
        order = np.argsort(ids, kind='stable')                 # updated_prob_rel sorts by index (ital.py:448)
        mi = 0.0
        for reli in itertools.product([False, True], repeat=D):                # ital.py:295
            pr = p_all[sum(int(r) << j for j, r in enumerate(reli))]
            log_pr = np.log(pr + EPS)
            for fbi in self._fb_iter(reli):                                     # ital.py:300-342
                if not any(fb != 0 for fb in fbi):
                    continue
                pr_upd = self._updated_prob_rel(reli, fbi, mean, cov, order)
                cur = np.log(pr_upd + EPS) - log_pr
                cur *= 1.0 if self._perfect_user() else self._likelihood(fbi, reli)   # ital.py:208
                if self.label_estimation == 'optimistic':                       # ital.py:210-219
                    if cur > mi:
                        mi = cur
                elif self.label_estimation == 'pessimistic':
                    if (mi == 0) or (cur < mi):
                        mi = cur
                else:
                    mi += cur * pr
        return mi

    def _fb_iter(self, reli):
        if self._perfect_user():
            return [[1 if r else -1 for r in reli]]
        elif self.label_prob >= 1:
            return itertools.product([-1, 1], repeat=len(reli))
        return itertools.product([-1, 0, 1], repeat=len(reli))

    def _likelihood(self, fbi, reli):                                           # ital.py:453-481
        prob = 1.0
        for fb, r in zip(fbi, reli):
            if fb == 0:
                prob *= 1.0 - self.label_prob
            elif fb == 2 * r - 1:
                prob *= self.label_prob * (1.0 - self.mistake_prob)
            else:
                prob *= self.label_prob * self.mistake_prob
        return prob

    def _updated_prob_rel(self, reli, fbi, mean, cov, order):
        """P(R = r | F = f) on the block (ital.py:432-450 -> retrieval_base.py:129-164 -> gp.py:295-344).

        The reference extends K^-1 by the annotated samples (extend_inv, gp.py:40-87) and predicts the block;
        that equals conditioning the block's current posterior N(mean, cov) on noisy observations of the
        annotated members, which is what is evaluated here.
        """
        fbi = np.asarray(fbi, dtype=np.float64)
        obs = np.nonzero(fbi != 0)[0]
        A = cov[np.ix_(obs, obs)] + self.noise * np.eye(len(obs))
        G = np.linalg.solve(A, cov[obs, :])                                      # (|O|, D)
        mean_u = mean + G.T @ (fbi[obs] - mean[obs])
        cov_u = cov - cov[:, obs] @ G
        mean_u, cov_u = mean_u[order], cov_u[np.ix_(order, order)]
        rel_sorted = np.asarray(reli)[order]
        p = orthant_prob_all(mean_u, cov_u, snq_order(len(mean) - 1) or None)
        return p[sum(int(r) << j for j, r in enumerate(rel_sorted))]
