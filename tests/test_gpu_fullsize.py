"""BASELINE.json's full size (n = 1M, d = 512, batch 4) on one GPU: the oracle cannot score a million candidates
in test time, so the checks are size-independent properties plus oracle comparisons on what is cheap at any n."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def big():
    import bench
    from ital_b200 import ITAL
    X, assign = bench.syn_block(0, 1000000, 512)
    learner = ITAL(X, length_scale=1.0)
    fbs = bench.labelled_state(assign[:65536])
    for fb in fbs:
        learner.update(fb)
    return X, assign, learner, fbs


def test_posterior_moments_match_oracle_on_a_sample(big):
    """rel_mean / variance of 4096 random rows against the oracle GP (gp.predict needs only the labelled rows)."""
    from oracle.ital_oracle import OracleGP
    X, assign, learner, fbs = big
    idx = [i for fb in fbs for i in fb]
    y = [v for fb in fbs for v in fb.values()]
    gp = OracleGP(X[idx].astype(np.float64), 1.0)
    gp.fit(list(range(len(idx))), y)
    rows = np.random.default_rng(1).choice(len(X), 4096, replace=False)
    mean, var = gp.predict(X[rows].astype(np.float64), cov_mode='diag')
    np.testing.assert_allclose(learner.rel_mean[rows], mean, rtol=1e-6, atol=1e-9)
    got_var = learner.gp.predict_stored(cov_mode='diag')[1][rows]
    np.testing.assert_allclose(got_var, var, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(learner.gp.predict(X[rows[:256]].astype(np.float64)), mean[:256], rtol=1e-6, atol=1e-9)


def test_first_step_is_the_binary_entropy_of_the_posterior(big):
    """Step 0 scores every candidate in closed form: score_i = H(Phi(m_i / sqrt(v_i))) (ital.py:364-369)."""
    from scipy.special import ndtr
    X, assign, learner, fbs = big
    learner.exhaustive = False
    ret = learner._fetch_stepwise(1, keep_scores=True)
    sc = learner.last_step_scores[0]
    seen = np.array([i for fb in fbs for i in fb])
    assert np.all(np.isnan(sc[seen])) and np.isnan(sc).sum() == len(seen)
    m = learner.rel_mean
    v = np.maximum(learner.gp.predict_stored(cov_mode='diag')[1], 0)
    z = m / np.sqrt(v)
    eps = 1e-12
    want = sum(p * (np.log(1 + eps) - np.log(p + eps)) for p in (ndtr(z), ndtr(-z)))
    ok = ~np.isnan(sc)
    np.testing.assert_allclose(sc[ok], want[ok], rtol=1e-6, atol=1e-12)
    assert ret[0] == int(np.nanargmax(sc))


def test_batch_is_invariant_under_pruning_and_lazy_rows(big):
    """Lazy-greedy pruning and on-demand projections are exact: same batch and bit-identical scores as scoring all
    10^6 candidates at every step; gains are non-increasing along the greedy path (submodularity)."""
    X, assign, learner, fbs = big
    learner.exhaustive, learner.lazy_rows = False, False
    base = learner.fetch_unlabelled(4)
    base_scores = np.array(learner.last_fetch_scores)
    assert len(set(base)) == 4 and not (set(base) & {i for fb in fbs for i in fb})
    learner.lazy_rows = None                      # the default: one persistent kernel per fetch
    assert learner.fetch_unlabelled(4) == base and learner.last_fused_steps == 4
    assert np.array_equal(np.array(learner.last_fetch_scores), base_scores)
    learner.lazy_rows, learner.fused = True, False      # the same as separate kernels
    assert learner.fetch_unlabelled(4) == base and learner.last_fused_steps == 0
    assert np.array_equal(np.array(learner.last_fetch_scores), base_scores)
    learner.fused = True
    learner.exhaustive = True
    assert learner._fetch_stepwise(4, keep_scores=True) == base
    np.testing.assert_allclose(learner.last_fetch_scores, base_scores, rtol=1e-12)
    full = learner.last_step_scores
    learner.lazy_rows = False
    assert learner.fetch_unlabelled(4) == base
    np.testing.assert_allclose(learner.last_fetch_scores, base_scores, rtol=1e-12)
    learner.exhaustive = False
    learner.lazy_rows = None
    # every candidate's conditional gain shrinks as the batch grows
    h = [0.0] + [float(s) for s in base_scores]
    gains = [full[t] - h[t] for t in range(4)]
    for t in range(1, 4):
        ok = ~np.isnan(gains[t])
        assert np.all(gains[t][ok] <= gains[t - 1][ok] + 1e-6)
    # the oracle re-scores the winners and their runners-up (a few rows are cheap at any n)
    from oracle.ital_oracle import OracleITAL
    top = sorted(set(base) | set(int(i) for t in range(4) for i in np.argsort(np.nan_to_num(full[t], nan=-1))[-5:])
                 | {i for fb in fbs for i in fb})
    pos = {g: k for k, g in enumerate(top)}
    ora = OracleITAL(X[top].astype(np.float64), length_scale=1.0)
    for fb in fbs:
        ora.update({pos[i]: v for i, v in fb.items()})
    ora.fetch_unlabelled(4, forced=[pos[i] for i in base])
    for t, tr in enumerate(ora.trace):
        glob = np.array(top)[tr['candidates']]
        np.testing.assert_allclose(full[t][glob], tr['scores'], rtol=1e-6, atol=1e-12)
        assert tr['argmax'] == pos[base[t]]


def test_top_results_is_the_sorted_ranking_at_full_size(big):
    """top_results() over 10^6 rows (245 tiles of the device radix sort): the permutation np.argsort would give, with
    exact ties in ascending row order (retrieval_base.py:64-75)."""
    X, assign, learner, fbs = big
    rm = np.array(learner.rel_mean)
    want = np.lexsort((np.arange(len(rm)), -rm))
    assert np.array_equal(learner.top_results(), want)
    assert np.array_equal(learner.top_results(100), want[:100])
