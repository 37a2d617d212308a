"""The oracle's fixed-node orthant rule (oracle/orthant.py) and the stand-in injected for the removed
scipy.stats.mvn.mvndst (oracle/ref_shims) against an INDEPENDENT implementation: scipy's own translation of Genz's
algorithm (scipy.stats.multivariate_normal.cdf, a randomised quasi-Monte-Carlo rule).  This un-circles the goldens
of three and more variables: they are produced by the unmodified reference driven through the stand-in, and the
stand-in is the oracle's rule at a higher order, so without this file the oracle would only be checked against itself
there.  (One and two variables are pinned to norm.cdf and Genz's BVU in tests/test_oracle_golden.py.)

The reference's call: mvndst(pivot, pivot, infin, correl, maxpts=100*dim, abseps=releps=1e-4), ital/ital.py:373-383.
"""
import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle.orthant import orthant_prob_all, snq_order
from oracle.ref_shims import STANDIN_Q, STANDIN_R, mvndst_standin


def _random_block(rng, D, strength):
    A = rng.standard_normal((D, D))
    cov = strength * (A @ A.T) / D + np.diag(rng.uniform(0.3, 1.0, D))
    mean = rng.uniform(-0.8, 0.8, D) * np.sqrt(np.diag(cov))
    return mean, cov


def _scipy_orthants(mean, cov, seed, maxpts=1000000):
    """All 2^D orthant probabilities by scipy's Genz QMC: P(sign pattern r) = P(S z <= 0), S = diag(-1 where r_j = 1)."""
    D = len(mean)
    out = np.empty(1 << D)
    for r in range(1 << D):
        sgn = np.array([-1.0 if (r >> j) & 1 else 1.0 for j in range(D)])
        mvn = multivariate_normal(mean=sgn * mean, cov=cov * np.outer(sgn, sgn), allow_singular=True,
                                  maxpts=maxpts, abseps=1e-10, releps=1e-10, seed=seed)
        out[r] = mvn.cdf(np.zeros(D))
    return out


# (variables, tolerance of the oracle's order, tolerance of the stand-in's order).  scipy's own QMC noise at 10^6 points
# is 1e-8 .. 1.5e-7 (measured; with 1.6 * 10^7 points the stand-in's order agrees with scipy to 4e-9 and the oracle's
# order to 1.4e-7, but that takes minutes per case)
CASES = [(3, 2e-7, 2e-7), (4, 5e-7, 5e-7)]


@pytest.mark.parametrize('D,tol_oracle,tol_standin', CASES)
def test_oracle_rule_matches_scipy_genz(D, tol_oracle, tol_standin):
    rng = np.random.default_rng(100 + D)
    for trial in range(4):
        mean, cov = _random_block(rng, D, strength=(0.2, 1.0, 2.0, 0.6)[trial])
        want = _scipy_orthants(mean, cov, seed=trial)
        assert abs(want.sum() - 1.0) < 1e-6
        got = orthant_prob_all(mean, cov, snq_order(D - 1) or None)          # the order the oracle scores with
        assert abs(got.sum() - 1.0) < 1e-7           # total mass of the rule (1e-11 at D = 3, 2e-8 at D = 4)
        assert np.max(np.abs(got - want)) <= tol_oracle, (D, trial, np.max(np.abs(got - want)))
        hi = orthant_prob_all(mean, cov, q=STANDIN_Q[D], R=STANDIN_R)        # the order the goldens were made with
        assert np.max(np.abs(hi - want)) <= tol_standin, (D, trial, np.max(np.abs(hi - want)))
        assert np.max(np.abs(hi - got)) <= 2e-7      # the two orders of the rule agree at the level of scipy's noise


@pytest.mark.parametrize('D', [3, 4])
def test_mvndst_standin_contract(D):
    """The stand-in answers the reference's exact call convention (standardised limits, INFIN, CORREL in
    np.tril_indices order) with the probability scipy's Genz rule gives."""
    rng = np.random.default_rng(7 + D)
    mean, cov = _random_block(rng, D, strength=1.0)
    stdev = np.sqrt(np.diag(cov))
    pivot = -mean / stdev
    i, j = np.tril_indices(D, -1)
    correl = cov[i, j] / (stdev[i] * stdev[j])
    want = _scipy_orthants(mean, cov, seed=3)
    for r in (0, (1 << D) - 1, 5 % (1 << D), 2):
        infin = np.array([(r >> k) & 1 for k in range(D)])
        err, pr, info = mvndst_standin(pivot, pivot, infin, correl, maxpts=D * 100, abseps=1e-4, releps=1e-4)
        assert info == 0 and abs(pr - want[r]) <= 1e-6


@pytest.mark.parametrize('D,tol', [(5, 2e-6), (6, 3e-6), (7, 5e-5)])
def test_rules_of_longer_batches_match_scipy_genz(D, tol):
    """Batches of 5, 6 and 7 samples (configs/toy.conf ships batch_size = 6): the tensor rule at 12 / 10 nodes per
    panel for 4 / 5 base variables and the sequential-conditioning lattice from 6 base variables on, against scipy's
    Genz QMC (2 * 10^5 points: noise ~5e-7).  With 1.6 * 10^7 points the measured differences are 8e-8 (D = 5), 5e-7
    (D = 6) and 9e-6 (D = 7) in the probabilities, 5e-7 / 4e-6 / 1e-4 in the entropy score."""
    rng = np.random.default_rng(100 + D)
    mean, cov = _random_block(rng, D, strength=0.6)
    want = _scipy_orthants(mean, cov, seed=D, maxpts=200000)
    got = orthant_prob_all(mean, cov, snq_order(D - 1) or None)
    assert abs(got.sum() - 1.0) < 1e-5
    assert np.max(np.abs(got - want)) <= tol, (D, np.max(np.abs(got - want)))


@pytest.mark.parametrize('D', [6, 8])
def test_single_orthant_lattice_matches_scipy_genz(D):
    """change_estimation_subset: the probability that batch + subset (+ candidate) land in ONE orthant of 6 to 9
    variables comes from 16 384 lattice nodes inside that orthant (oracle/orthant.py sc_orthant = csrc/snq_host.h
    sc_orthant) with the last variable analytic (oracle/ce_subset.py orthant_prob_any).  Against scipy's Genz rule at
    2 * 10^6 points: 1e-4 relative for orthants that carry at least a percent of the mass -- the accuracy class stated
    for this mode (the reference's own mvndst(maxpts = 100 * dim) is 1e-3 there)."""
    from oracle.ce_subset import SUB_N, orthant_prob_any
    rng = np.random.default_rng(300 + D)
    for trial in range(2):
        mean, cov = _random_block(rng, D, strength=(0.3, 1.0)[trial])
        rel = mean > 0                                   # the most likely orthant
        sgn = np.where(rel, -1.0, 1.0)
        want = multivariate_normal(mean=sgn * mean, cov=cov * np.outer(sgn, sgn), allow_singular=True,
                                   maxpts=2000000, abseps=1e-10, releps=1e-10, seed=trial).cdf(np.zeros(D))
        got = orthant_prob_any(rel, mean, cov, n_lattice=SUB_N)
        assert want > 1e-2
        assert abs(got - want) <= 1e-4 * want, (D, trial, got, want)
