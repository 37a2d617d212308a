"""Parity of the CUDA path (through the C ABI) with the oracle and with the goldens recorded from the reference.

Bar (SURVEY.md 8c / BASELINE.json north_star): selected indices identical (exact ties resolved to the lowest
index; where the reference's own maximum is an exact floating-point tie the recorded choice must score within
1e-9 relative of the GPU's maximum); rel_mean, variances and MI scores within 1e-6 relative of the float64
oracle (absolute floor 1e-9 for moments, 1e-12 = the reference's eps for scores).
"""
import numpy as np
import pytest

from conftest import (baseline_names, clip_names, drive, golden_names, load_baseline, load_clip, load_golden,
                      load_subset, load_updpred, subset_names, updpred_names)

pytestmark = pytest.mark.gpu

SCORE_RTOL, SCORE_ATOL = 1e-6, 1e-12


def _gpu_learner(X, **kw):
    from ital_b200 import ITAL
    return ITAL(X, **kw)


def _perfect(g):
    return float(g['label_prob']) >= 1 and float(g['mistake_prob']) <= 0 and str(g['label_estimation']) == 'mean'


PERFECT = [n for n in golden_names() if _perfect(load_golden(n))]


def _compare_steps(gpu, ora, k):
    """Exhaustive GPU scores of every greedy step against the oracle re-scored along the GPU's own path."""
    ret = gpu._fetch_stepwise(k, keep_scores=True)
    ora.fetch_unlabelled(k, forced=ret)
    assert len(gpu.last_step_scores) == len(ora.trace) == len(ret)
    kinds = []
    for t, (sc, tr) in enumerate(zip(gpu.last_step_scores, ora.trace)):
        cand = tr['candidates']
        got = sc[cand]
        assert not np.any(np.isnan(got)), 'step %d: unscored candidates' % t
        others = np.setdiff1d(np.arange(len(sc)), cand)
        assert np.all(np.isnan(sc[others]))
        np.testing.assert_allclose(got, tr['scores'], rtol=SCORE_RTOL, atol=SCORE_ATOL, err_msg='step %d' % t)
        if tr['argmax'] == ret[t]:
            kinds.append('identical')
        else:       # permitted only for an exact tie in the oracle's own scores
            pos = int(np.nonzero(cand == ret[t])[0][0])
            assert tr['scores'][pos] >= tr['scores'].max() * (1 - 1e-12), (t, ret[t], tr['argmax'])
            kinds.append('tie')
    return ret, kinds


@pytest.mark.parametrize('name', PERFECT)
def test_golden_parity(name):
    from oracle.ital_oracle import OracleITAL
    g = load_golden(name)
    kw = dict(g['learner_kw'])
    for storage in ('float64',):
        gpu = drive(_gpu_learner(g['X'], queries=list(g['queries']), storage=storage, exhaustive=True, **kw), g)
        ora = drive(OracleITAL(g['X'], queries=list(g['queries']), **kw), g)
        np.testing.assert_allclose(gpu.rel_mean, g['rel_mean'], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(gpu.rel_mean, ora.rel_mean, rtol=1e-6, atol=1e-9)
        var = gpu.gp.predict_stored(cov_mode='diag')[1][:len(g['X'])]
        np.testing.assert_allclose(var, g['var_diag'], rtol=1e-6, atol=1e-9 * float(g['var']))
        if kw.get('top_candidates') is not None:
            ret = gpu.fetch_unlabelled(int(g['k']))
        else:
            ret, kinds = _compare_steps(gpu, ora, int(g['k']))
        # against the reference's own recorded choices
        for t, st in enumerate(g['steps']):
            if ret[t] == st['chosen']:
                continue
            pos = int(np.nonzero(st['candidates'] == ret[t])[0][0])
            assert st['mi'][pos] >= st['mi'].max() * (1 - 1e-9), (name, t, ret[t], st['chosen'])
            break       # after an exact tie the two greedy paths are different batches
        # lazy-greedy pruning must not change the batch
        gpu.exhaustive = False
        assert gpu.fetch_unlabelled(int(g['k'])) == ret
        if kw.get('top_candidates') is None:      # (the restriction is applied by fetch_unlabelled itself)
            assert gpu._fetch_stepwise(int(g['k'])) == ret


@pytest.mark.parametrize('name', [n for n in PERFECT if load_golden(n).get('top_candidates', -1) < 0])
def test_batch_covariances_match_the_reference(name):
    """predict_cov_batch (ital/gp.py:235-261; AppendedMutualInformation.append, ital/ital.py:561-586): the reference's
    N x (t+1) x (t+1) covariances of ret + [i] are never stored here -- every row keeps one more projection entry per
    selected point (the incremental Cholesky row l_i).  Rebuilt from the point records exported in the middle of a
    fetch: C_base = L_b L_b^T from the selected rows' own entries, c_i = L_b l_i, var_i from the record header."""
    g = load_golden(name)
    kw = dict(g['learner_kw'])
    gpu = drive(_gpu_learner(g['X'], queries=list(g['queries']), storage='float64', exhaustive=True, lazy_rows=False,
                             **kw), g)
    sh = gpu._shard
    W = int(sh.lib.ital_width(sh.handle))
    sh.fetch_begin(kw['label_prob'], kw['mistake_prob'])
    try:
        ret = []
        for t, st in enumerate(g['steps']):
            cand = [int(i) for i in st['candidates']]
            recs = np.concatenate([sh.export_points(ret + cand[lo:lo + 48]) for lo in range(0, len(cand), 48)]) \
                if t else sh.export_points(cand[:1])[:0]
            h = 8
            if t == 0:
                var = np.concatenate([sh.export_points(cand[lo:lo + 64])[:, 5] for lo in range(0, len(cand), 64)])
                np.testing.assert_allclose(np.maximum(var, 0)[:, None, None], st['rel_covs'], rtol=1e-6, atol=1e-9)
            else:
                pos = 0
                for lo in range(0, len(cand), 48):
                    nblk = len(cand[lo:lo + 48])
                    blk = recs[pos:pos + t + nblk]
                    pos += t + nblk
                    Lb = np.tril(blk[:t, h + W:h + W + t])
                    li = blk[t:, h + W:h + W + t]
                    cov = np.empty((nblk, t + 1, t + 1))
                    cov[:, :t, :t] = Lb @ Lb.T
                    cov[:, :t, t] = cov[:, t, :t] = li @ Lb.T
                    cov[:, t, t] = blk[t:, 5]
                    np.testing.assert_allclose(cov, st['rel_covs'][lo:lo + nblk], rtol=1e-6, atol=1e-9,
                                               err_msg='%s step %d' % (name, t))
                    # the conditional variance the scorer uses: v_i - |l_i|^2 (record header [3])
                    np.testing.assert_allclose(blk[t:, 3], blk[t:, 5] - (li * li).sum(axis=1), rtol=1e-9, atol=1e-12)
            rec = sh.fetch_propose(-np.inf, True)
            if int(rec[0]) != st['chosen']:
                break                               # an exact tie in the reference's own scores: different batch from here
            ret.append(int(rec[0]))
            if t + 1 < len(g['steps']):
                sh.fetch_commit(rec)
    finally:
        sh.fetch_end()


def test_mistaken_user_golden_and_oracle():
    """label_prob = 1, mistake_prob = 0.5 (configs/butterflies-aggressive.conf): the scores are the perfect-user
    scores plus a per-step constant.  Against the oracle's evaluation of the same sum (oracle/general_sets.py) at 1e-6;
    against the reference golden and the oracle's literal double loop at 1e-4 relative (both carry log(p' + 1e-12)
    terms with p' at round-off level, see tests/test_oracle_golden.py)."""
    from oracle.ital_oracle import OracleITAL
    g = load_golden('butterflies_aggressive_k3')
    kw = dict(g['learner_kw'])
    gpu = drive(_gpu_learner(g['X'], exhaustive=True, **kw), g)
    ora = drive(OracleITAL(g['X'], general_sets=True, **kw), g)
    lit = drive(OracleITAL(g['X'], **kw), g)
    ret = gpu._fetch_stepwise(int(g['k']), keep_scores=True)
    assert ret == g['ret'].tolist()
    ora.fetch_unlabelled(int(g['k']), forced=ret)
    lit.fetch_unlabelled(int(g['k']), forced=ret)
    for t, (sc, tr, tl, st) in enumerate(zip(gpu.last_step_scores, ora.trace, lit.trace, g['steps'])):
        np.testing.assert_allclose(sc[tr['candidates']], tr['scores'], rtol=SCORE_RTOL, atol=SCORE_ATOL)
        np.testing.assert_allclose(sc[tl['candidates']], tl['scores'], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sc[st['candidates']], st['mi'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gpu.last_fetch_scores, [st['mi'].max() for st in g['steps']], rtol=1e-4)
    gpu.exhaustive = False
    assert gpu.fetch_unlabelled(int(g['k'])) == ret
    assert gpu.last_fused_steps == int(g['k'])           # (the constant is added inside the persistent kernel too)
    np.testing.assert_allclose(gpu.last_fetch_scores, [tr['scores'].max() for tr in ora.trace], rtol=SCORE_RTOL)


@pytest.mark.parametrize('name', ['toy_mistakes_k3', 'butterflies_conservative_k3'])
def test_general_feedback_model_golden_and_oracle(name):
    """label_prob < 1 (configs/toy-mistakes.conf, configs/butterflies-conservative.conf): conditional node sets,
    every candidate scored.  1e-6 relative against the oracle's evaluation with the same shared conditional node sets
    (oracle/general_sets.py); 1e-5 against the oracle's literal enumeration and 1e-4 against the reference golden
    (different node placement for the same integrals)."""
    from oracle.ital_oracle import OracleITAL
    g = load_golden(name)
    kw = dict(g['learner_kw'])
    gpu = drive(_gpu_learner(g['X'], **kw), g)
    ora = drive(OracleITAL(g['X'], general_sets=True, **kw), g)
    lit = drive(OracleITAL(g['X'], **kw), g)
    ret = gpu._fetch_stepwise(int(g['k']), keep_scores=True)
    ora.fetch_unlabelled(int(g['k']), forced=ret)
    lit.fetch_unlabelled(int(g['k']), forced=ret)
    for t, (sc, tr, tl) in enumerate(zip(gpu.last_step_scores, ora.trace, lit.trace)):
        np.testing.assert_allclose(sc[tr['candidates']], tr['scores'], rtol=SCORE_RTOL, atol=1e-9, err_msg='step %d' % t)
        # (the same candidate, or one the oracle scores within round-off of its maximum: symmetric rows of the toy set)
        assert tr['argmax'] == ret[t] or tr['scores'][list(tr['candidates']).index(ret[t])] >= \
            tr['scores'].max() - 1e-9 * abs(tr['scores'].max())
        np.testing.assert_allclose(sc[tl['candidates']], tl['scores'], rtol=1e-5, atol=1e-9, err_msg='step %d' % t)
    for t, st in enumerate(g['steps']):          # the reference's own record, as far as the paths coincide
        if ret[t] != st['chosen']:
            pos = int(np.nonzero(st['candidates'] == ret[t])[0][0])
            assert st['mi'][pos] >= st['mi'].max() - 1e-4 * abs(st['mi'].max())
            break
        np.testing.assert_allclose(gpu.last_step_scores[t][st['candidates']], st['mi'], rtol=1e-4, atol=1e-6)
    assert gpu.fetch_unlabelled(int(g['k'])) == ret


def _syn(n, d, seed=0, centres=50):
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((centres, d))
    assign = rng.integers(0, centres, n)
    X = C[assign] + 0.6 * rng.standard_normal((n, d))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    return X.astype(np.float32).astype(np.float64), assign


def _label_syn(learner, assign):
    pos = np.nonzero(assign == assign[0])[0]
    neg = np.nonzero(assign != assign[0])[0]
    learner.update({0: 1})
    learner.update({**{int(i): 1 for i in pos[1:4]}, **{int(i): -1 for i in neg[:5]}})


@pytest.mark.parametrize('n,d,storage', [(3000, 512, 'auto'), (2500, 200, 'float64'), (1111, 70, 'float32'),
                                          (4097, 1000, 'auto')])
def test_synthetic_parity_exhaustive_and_pruned(n, d, storage):
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(n, d, seed=n)
    gpu = _gpu_learner(X, length_scale=1.0, storage=storage, exhaustive=True)
    ora = OracleITAL(X, length_scale=1.0)
    assert gpu.storage == ('float32' if storage in ('auto', 'float32') else 'float64')
    _label_syn(gpu, assign)
    _label_syn(ora, assign)
    np.testing.assert_allclose(gpu.rel_mean, ora.rel_mean, rtol=1e-6, atol=1e-9)
    ret, kinds = _compare_steps(gpu, ora, 4)
    assert kinds == ['identical'] * 4
    assert ret == ora.fetch_unlabelled(4)
    gpu.exhaustive = False
    # streaming passes (pipelined multi-kernel loop), then the same loop step by step
    gpu.lazy_rows = False
    assert gpu.fetch_unlabelled(4) == ret and gpu.last_fused_steps == 0
    stream_scores = np.array(gpu.last_fetch_scores)
    stats = gpu._fetch_stepwise(4) and gpu.last_fetch_stats
    assert np.array_equal(np.array(gpu.last_fetch_scores), stream_scores)
    # the default: projections on demand for the scored rows only, the whole fetch in one persistent kernel --
    # same batch, same scores bit for bit
    gpu.lazy_rows = None
    assert gpu.fetch_unlabelled(4) == ret and gpu.last_fused_steps == 4
    assert np.array_equal(np.array(gpu.last_fetch_scores), stream_scores)
    assert gpu.fetch_unlabelled(2) == ret[:2] and gpu.last_fused_steps == 2
    assert gpu.fetch_unlabelled(1) == ret[:1] and gpu.last_fused_steps == 1
    # ... and as separate kernels (ITAL_B200_FUSED=0)
    gpu.fused = False
    assert gpu.fetch_unlabelled(4) == ret and gpu.last_fused_steps == 0
    assert np.array_equal(np.array(gpu.last_fetch_scores), stream_scores)
    gpu.fused = True
    gpu.lazy_rows = True
    gpu.exhaustive = True
    assert gpu._fetch_stepwise(4, keep_scores=True) == ret
    ora.fetch_unlabelled(4, forced=ret)
    for sc, tr in zip(gpu.last_step_scores, ora.trace):
        np.testing.assert_allclose(sc[tr['candidates']], tr['scores'], rtol=SCORE_RTOL, atol=SCORE_ATOL)
    gpu.exhaustive = False
    gpu.lazy_rows = None
    # [0] rows in the final worklist, [1] rows scored by quadrature (-1: closed form), [2] nodes
    assert stats[0][1] == -1 and all(0 < s[1] <= n + 2 * 148 for s in stats[1:]), stats
    assert [int(s[2]) for s in stats] == [1, 64, 1024, 13824]


def test_bulk_staged_stream_is_bit_identical():
    """k_extend_bulk (TMA-staged rows) against k_extend (register loads): same bits in rel_mean and scores."""
    X, assign = _syn(20011, 512, seed=4)
    out = {}
    for bulk in (False, True):
        gpu = _gpu_learner(X, length_scale=1.0, bulk_stream=bulk, lazy_rows=False)
        _label_syn(gpu, assign)
        out[bulk] = (gpu.fetch_unlabelled(4), np.array(gpu.last_fetch_scores), gpu.rel_mean.copy())
    assert out[False][0] == out[True][0]
    assert np.array_equal(out[False][1], out[True][1]) and np.array_equal(out[False][2], out[True][2])


def test_batches_beyond_four_track_the_oracle():
    """Greedy steps with 4 and 5 base variables use the tensor rule at 12 / 10 nodes per panel (81k / 400k nodes, generated
    on the host), from 6 base variables on the sequential-conditioning lattice (oracle/orthant.py sc_nodes,
    csrc/snq_host.h generate_sc); same nodes on both sides, so scores agree to round-off.  The persistent kernel runs the
    first four steps and hands over to the multi-kernel loop; the lazy-greedy margin grows with the step's quadrature
    error (prune_margin)."""
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(500, 48, seed=21, centres=8)
    gpu = _gpu_learner(X, length_scale=1.0, exhaustive=True)
    ora = OracleITAL(X, length_scale=1.0)
    _label_syn(gpu, assign)
    _label_syn(ora, assign)
    ret, kinds = _compare_steps(gpu, ora, 8)
    assert len(ret) == 8 and len(set(ret)) == 8
    gpu.exhaustive = False
    assert gpu.fetch_unlabelled(8) == ret and gpu.last_fused_steps == 4
    gpu.lazy_rows = False
    assert gpu.fetch_unlabelled(8) == ret and gpu.last_fused_steps == 0
    with pytest.raises(NotImplementedError):
        gpu.fetch_unlabelled(12)


@pytest.mark.parametrize('var,noise,ls', [(2.5, 1e-4, 0.7), (0.3, 1e-6, 2.0), (1.0, 1e-2, 1.0)])
def test_kernel_hyperparameters(var, noise, ls):
    """var, sigma_noise and length_scale away from the defaults (ital/gp.py:100: k = var exp(-d^2 / 2 sigma^2))."""
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(1500, 40, seed=31, centres=9)
    kw = dict(length_scale=ls, var=var, noise=noise)
    gpu, ora = _gpu_learner(X, exhaustive=True, **kw), OracleITAL(X, **kw)
    for L in (gpu, ora):
        _label_syn(L, assign)
    np.testing.assert_allclose(gpu.rel_mean, ora.rel_mean, rtol=1e-6, atol=1e-9)
    v = gpu.gp.predict_stored(cov_mode='diag')[1]
    np.testing.assert_allclose(v, ora.gp.predict_stored(cov_mode='diag')[1], rtol=1e-6, atol=1e-9 * var)
    ret, kinds = _compare_steps(gpu, ora, 4)
    assert ret == ora.fetch_unlabelled(4)
    Xt = X[:64] * 1.01
    np.testing.assert_allclose(gpu.gp.predict(Xt), ora.gp.predict(Xt), rtol=1e-6, atol=1e-9)


def test_entropy_sampling_shares_the_kernels():
    """EntropySampling (ital/baseline_methods.py:229-287): joint entropy of the batch = perfect-user ITAL."""
    from ital_b200 import EntropySampling
    X, assign = _syn(2000, 64, seed=17, centres=10)
    a, b = EntropySampling(X, length_scale=1.0, mistake_prob=0.3), _gpu_learner(X, length_scale=1.0)
    for L in (a, b):
        _label_syn(L, assign)
    assert a.fetch_unlabelled(4) == b.fetch_unlabelled(4)
    assert np.array_equal(a.last_fetch_scores, b.last_fetch_scores)


@pytest.mark.parametrize('name', baseline_names())
def test_entropy_sampling_matches_the_reference(name):
    """EntropySampling.fetch_unlabelled of the UNMODIFIED reference (ital/baseline_methods.py:241-287: single_entropy
    for the first sample, batch_entropy = -sum p log p over the joint sign patterns afterwards): same batch, and every
    candidate's entropy of every greedy step within 1e-6 (the reference clips single probabilities to [1e-8, 1 - 1e-8]
    and drops terms below 1e-12; ITAL's summand is p (log(1 + 1e-12) - log(p + 1e-12)): differences below 1e-10)."""
    from ital_b200 import EntropySampling
    g = load_baseline(name)
    gpu = drive(EntropySampling(g['X'], exhaustive=True, **g['learner_kw']), g)
    ret = gpu._fetch_stepwise(int(g['k']), keep_scores=True)
    assert ret == g['entropy_ret'].tolist()
    for t, (sc, st) in enumerate(zip(gpu.last_step_scores, g['entropy_steps'])):
        # step 0: the reference's clip of p to [1e-8, 1 - 1e-8] lifts the entropy of decided samples to 1.9e-7
        atol = 2.5e-7 if t == 0 else (2e-6 if t >= 2 else 1e-9)
        np.testing.assert_allclose(sc[st['candidates']], st['entropy'], rtol=1e-6, atol=atol, err_msg='%s step %d' % (name, t))
    gpu.exhaustive = False
    assert gpu.fetch_unlabelled(int(g['k'])) == ret


@pytest.mark.parametrize('name', baseline_names())
def test_variance_sampling_matches_the_reference(name):
    """VarianceSampling.fetch_unlabelled of the UNMODIFIED reference (ital/baseline_methods.py:110-155), without and with
    use_correlations (greedy on sum of variances - sum of covariances, from predict_cov_batch there, from the
    incremental Cholesky rows here)."""
    from ital_b200 import VarianceSampling
    g = load_baseline(name)
    plain = drive(VarianceSampling(g['X'], **g['learner_kw']), g)
    np.testing.assert_allclose(np.maximum(plain.gp.predict_stored(cov_mode='diag')[1], 0), g['var_diag'], rtol=1e-6, atol=1e-9)
    assert plain.fetch_unlabelled(int(g['k'])) == g['variance_ret'].tolist()
    corr = drive(VarianceSampling(g['X'], use_correlations=True, **g['learner_kw']), g)
    assert corr.fetch_unlabelled(int(g['k'])) == g['variance_corr_ret'].tolist()
    assert corr.fetch_unlabelled(2) == g['variance_corr_ret'].tolist()[:2]          # (nothing of a fetch persists)


def test_repeated_rounds_track_the_oracle():
    """Several update/fetch rounds like run_experiment.py:160-164, incremental model on the GPU."""
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(1500, 64, seed=7, centres=12)
    y = np.where(assign == assign[0], 1, -1)
    gpu = _gpu_learner(X, length_scale=0.9)
    ora = OracleITAL(X, length_scale=0.9)
    for L in (gpu, ora):
        L.update({0: 1})
    for rnd in range(5):
        gpu.lazy_rows = (False, True, None)[rnd % 3]  # streaming passes, on-demand projections, the default (fused)
        gpu.fused = rnd != 1
        a, b = gpu.fetch_unlabelled(4), ora.fetch_unlabelled(4)
        assert a == b, rnd
        fb = {i: int(y[i]) for i in a}
        gpu.update(fb)
        ora.update(fb)
        np.testing.assert_allclose(gpu.rel_mean, ora.rel_mean, rtol=1e-6, atol=1e-9)
        assert np.array_equal(gpu.top_results(10), ora.top_results(10))
    assert gpu.rounds == ora.rounds == 6
    Xt = X[::7] + 0.01
    np.testing.assert_allclose(gpu.gp.predict(Xt), ora.gp.predict(Xt), rtol=1e-6, atol=1e-9)
    m, v = gpu.gp.predict(Xt, cov_mode='diag')
    mo, vo = ora.gp.predict(Xt, cov_mode='diag')
    np.testing.assert_allclose(v, vo, rtol=1e-6, atol=1e-9)


def test_interface_edge_cases():
    X, assign = _syn(12, 8, seed=3, centres=4)
    gpu = _gpu_learner(X, length_scale=1.0)
    assert gpu.rel_mean is None
    with pytest.raises(RuntimeError):
        gpu.fetch_unlabelled(2)                       # nothing labelled yet
    gpu.update({3: 1, 5: -1, 9: 0})
    assert gpu.rounds == 1 and gpu.relevant_ids == {3} and gpu.irrelevant_ids == {5} and gpu.unnameable_ids == {9}
    with pytest.raises(RuntimeError, match='Cannot change feedback once given.'):
        gpu.update({3: -1})
    gpu.update({3: 1})                                # same label again: silently ignored
    assert gpu.rounds == 1
    gpu.update({11: 0})                               # only unnameable: no round counted
    assert gpu.rounds == 1
    ret = gpu.fetch_unlabelled(100)                   # k clamped to the unseen rows (ital.py:99-100)
    assert sorted(ret) == sorted(set(range(12)) - {3, 5, 9, 11})
    assert gpu.get_unseen() == sorted(set(range(12)) - {3, 5, 9, 11})
    assert gpu.fetch_unlabelled(0) == []
    gpu.reset()
    assert gpu.rel_mean is None and gpu.rounds == 0 and gpu.get_unseen() == list(range(12))
    for kw in (dict(change_estimation_subset=4, label_prob=0.5), dict(change_estimation_subset=None)):
        bad = _gpu_learner(X, length_scale=1.0, **kw)
        bad.update({0: 1})
        with pytest.raises(NotImplementedError):
            bad.fetch_unlabelled(2)


def test_duplicate_rows_tie_to_lowest_index():
    X, assign = _syn(300, 32, seed=11, centres=6)
    X[200] = X[17]
    X[250] = X[17]
    from oracle.ital_oracle import OracleITAL
    gpu, ora = _gpu_learner(X, length_scale=1.0), OracleITAL(X, length_scale=1.0)
    for L in (gpu, ora):
        L.update({0: 1, 100: -1})
    a, b = gpu.fetch_unlabelled(4), ora.fetch_unlabelled(4)
    assert a == b
    assert not ({200, 250} & set(a)) or 17 in a


@pytest.mark.parametrize('n', [1, 31, 4096, 4097, 70001])
def test_top_results_sorted_on_the_device(n):
    """retrieval_base.py:64-75: np.argsort(rel_mean)[::-1][:k]; ties (duplicate rows) in ascending row order;
    tile boundaries of the radix sort (4096 keys per block); query rows are never returned."""
    X, assign = _syn(n, 16, seed=n, centres=5)
    if n > 40:
        X[n // 2] = X[3]
        X[n - 1] = X[3]                               # exact ties
    queries = [X[0] + 0.01, X[n // 3] - 0.01]
    gpu = _gpu_learner(X, queries=queries, length_scale=1.0)
    if n > 2:
        gpu.update({1: -1, 2: 1})
    rm = np.array(gpu.rel_mean)
    assert len(rm) == n
    want = np.lexsort((np.arange(n), -rm))
    assert np.array_equal(gpu.top_results(), want)
    for k in (1, 7, n // 2, n, n + 5):
        got = gpu.top_results(k)
        assert got.dtype == np.int64 and np.array_equal(got, want[:k])
    assert len(gpu.top_results(0)) == 0
    assert np.all(np.diff(rm[gpu.top_results()]) <= 0)
    assert np.array_equal(np.sort(rm)[::-1], rm[gpu.top_results()])       # same values as the reference's argsort


@pytest.mark.parametrize('name', updpred_names())
def test_updated_prediction_matches_reference(name):
    """ActiveRetrievalBase.updated_prediction (retrieval_base.py:129-164, gp.py:295-344) against the outputs of the
    unmodified reference: mean, 'diag' variance and 'full' covariance; the model itself must not change."""
    g = load_updpred(name)
    gpu = _gpu_learner(g['X'], queries=list(g['queries']), **g['learner_kw'])
    for fb in g['updates']:
        gpu.update(fb)
    before = np.array(gpu.rel_mean)
    for pr in g['probes']:
        np.testing.assert_allclose(gpu.updated_prediction(pr['feedback'], pr['test'], cov_mode=None), pr['mean'],
                                   rtol=1e-6, atol=1e-9)
        m, v = gpu.updated_prediction(pr['feedback'], pr['test'], cov_mode='diag')
        np.testing.assert_allclose(v, pr['var'], rtol=1e-6, atol=1e-9)
        m, c = gpu.updated_prediction(pr['feedback'], pr['test'])
        np.testing.assert_allclose(m, pr['mean'], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(c, pr['cov'], rtol=1e-6, atol=1e-9)
    m, c = gpu.gp.predict_stored(g['probes'][0]['test'], cov_mode='full')
    m2, c2 = gpu.updated_prediction({}, g['probes'][0]['test'])
    assert np.array_equal(m, m2) and np.array_equal(c, c2)
    assert np.array_equal(before, gpu.rel_mean)
    with pytest.raises(RuntimeError, match='Cannot change feedback once given.'):
        first = next(iter(g['updates'][0].items()))
        gpu.updated_prediction({first[0]: -first[1]}, g['probes'][0]['test'])


@pytest.mark.parametrize('n,d', [(2053, 512), (999, 400), (64, 512)])
def test_multi_label_update_on_2kb_rows(n, d):
    """GaussianProcess.update with 2, 3 and 4 samples at once (gp.py:164-200) on float32 rows of 2 KB, where the
    pass runs on the bulk-copy ring with four-row slots (k_extend_bulk_multi); ragged tails (n not a multiple of 4)."""
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(n, d, seed=n + d, centres=9)
    y = np.where(assign == assign[0], 1, -1)
    gpu = _gpu_learner(X, length_scale=1.0, storage='float32')
    ora = OracleITAL(X, length_scale=1.0)
    nxt = 0
    for q in (1, 2, 3, 4, 4, 3):
        fb = {i: int(y[i]) for i in range(nxt, nxt + q)}
        nxt += q
        gpu.update(fb)
        ora.update(fb)
        np.testing.assert_allclose(gpu.rel_mean, ora.rel_mean, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(gpu.gp.predict_stored(cov_mode='diag')[1],
                                   ora.gp.predict_stored(cov_mode='diag')[1][:n], rtol=1e-6, atol=1e-9)
    assert gpu.fetch_unlabelled(3) == ora.fetch_unlabelled(3)


def test_many_labelled_points_need_more_than_48kb_of_shared_memory():
    """More than ~500 labelled points double the projection capacity to 1024 columns: k_catchup then asks for 64 KB
    of dynamic shared memory (opt-in above 48 KB), and the persistent fetch kernel hands over to the multi-kernel loop
    (its staging would not fit); gp.predict with more than 1536 labelled points needs the same opt-in."""
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(900, 24, seed=77, centres=7)
    y = np.where(assign == assign[0], 1, -1)
    gpu, ora = _gpu_learner(X, length_scale=1.2, noise=1e-4), OracleITAL(X, length_scale=1.2, noise=1e-4)
    lab = {i: int(y[i]) for i in range(0, 520)}
    for L in (gpu, ora):
        L.update({0: 1})
        L.update({i: v for i, v in lab.items() if i != 0})
    assert int(gpu._shard.lib.ital_width_cap(gpu._shard.handle)) >= 1024
    np.testing.assert_allclose(gpu.rel_mean, ora.rel_mean, rtol=1e-6, atol=1e-8)
    a, b = gpu.fetch_unlabelled(3), ora.fetch_unlabelled(3)
    assert a == b and gpu.last_fused_steps == 0
    np.testing.assert_allclose(gpu.last_fetch_scores, [t['scores'].max() for t in ora.trace], rtol=1e-6, atol=1e-9)
    gpu.lazy_rows = False
    assert gpu.fetch_unlabelled(3) == a
    Xt = X[::9] * 0.99
    np.testing.assert_allclose(gpu.gp.predict(Xt), ora.gp.predict(Xt), rtol=1e-6, atol=1e-8)


def test_predict_with_more_than_1536_labelled_points():
    from oracle.ital_oracle import OracleITAL
    rng = np.random.default_rng(3)
    X = rng.standard_normal((1700, 6))
    y = np.sign(X[:, 0] + 0.1)
    gpu, ora = _gpu_learner(X, length_scale=0.8, noise=1e-3), OracleITAL(X, length_scale=0.8, noise=1e-3)
    fb = {i: int(y[i]) for i in range(1600)}
    for L in (gpu, ora):
        L.update(fb)
    Xt = rng.standard_normal((40, 6))
    m, v = gpu.gp.predict(Xt, cov_mode='diag')
    mo, vo = ora.gp.predict(Xt, cov_mode='diag')
    np.testing.assert_allclose(m, mo, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(v, vo, rtol=1e-5, atol=1e-7)


def test_interface_errors_before_any_work():
    X, assign = _syn(60, 8, seed=13, centres=4)
    gpu = _gpu_learner(X, length_scale=1.0, label_prob=0.75, mistake_prob=0.2)
    gpu.update({0: 1})
    with pytest.raises(NotImplementedError, match='label_prob < 1'):
        gpu.fetch_unlabelled(6)                       # configs/toy-mistakes.conf ships batch_size = 6
    assert len(gpu.fetch_unlabelled(2)) == 2
    for bad in (60, -1):
        with pytest.raises(IndexError):
            gpu.update({bad: 1})
    with _gpu_learner(X, length_scale=1.0) as ctx:    # context manager releases the GPU state
        ctx.update({1: 1})
        assert len(ctx.fetch_unlabelled(1)) == 1
    assert ctx._shard is None


def test_predict_full_covariance_and_device_top_candidates():
    """GaussianProcess.predict(cov_mode='full') (gp.py:285-287) from the test rows' projections; the top_candidates
    restriction (ital.py:111-117) computed on the device (masked sort of the means) against the oracle's argpartition."""
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(3000, 40, seed=41, centres=9)
    gpu, ora = _gpu_learner(X, length_scale=1.0), OracleITAL(X, length_scale=1.0)
    for L in (gpu, ora):
        _label_syn(L, assign)
    Xt = X[5:45] * 1.02
    m, c = gpu.gp.predict(Xt, cov_mode='full')
    mo, co = ora.gp.predict(Xt, cov_mode='full')
    np.testing.assert_allclose(m, mo, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(c, co, rtol=1e-6, atol=1e-9)
    for top in (50, 7.0):
        gpu.top_candidates = ora.top_candidates = top
        a, b = gpu.fetch_unlabelled(4), ora.fetch_unlabelled(4)
        assert a == b, (top, a, b)
        assert gpu.fetch_unlabelled(4) == a            # the restriction is lifted after every fetch and rebuilt
    gpu.top_candidates = ora.top_candidates = None
    assert gpu.fetch_unlabelled(3) == ora.fetch_unlabelled(3)


@pytest.mark.parametrize('est,lp,mp', [('optimistic', 1.0, 0.2), ('pessimistic', 1.0, 0.2), ('optimistic', 1.0, 0.0),
                                       ('pessimistic', 1.0, 0.0), ('optimistic', 0.6, 0.1), ('pessimistic', 0.6, 0.1),
                                       ('pessimistic', 0.5, 0.0)])
def test_label_estimation_optimistic_and_pessimistic(est, lp, mp):
    """ITAL(label_estimation='optimistic' | 'pessimistic') (ital/ital.py:210-215): the largest / the "first or smaller"
    single term of the enumeration over relevance and feedback configurations, in the reference's order (including the
    fold's quirk that a running value of exactly 0 -- a feedback configuration of likelihood 0 -- is replaced by the next
    term).  Every candidate of every step against the oracle's literal loop."""
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(70, 12, seed=5, centres=4)
    kw = dict(length_scale=1.0, label_prob=lp, mistake_prob=mp, label_estimation=est)
    gpu, ora = _gpu_learner(X, **kw), OracleITAL(X, **kw)
    for L in (gpu, ora):
        L.update({0: 1, 1: -1 if assign[1] != assign[0] else 1, 2: -1 if assign[2] != assign[0] else 1})
    ret = gpu._fetch_stepwise(3, keep_scores=True)
    ora.fetch_unlabelled(3, forced=ret)
    for t, (sc, tr) in enumerate(zip(gpu.last_step_scores, ora.trace)):
        # (a term is a weighted difference of two logarithms; where they nearly cancel its absolute accuracy is that of
        # the probabilities, ~1e-7, whatever its size)
        np.testing.assert_allclose(sc[tr['candidates']], tr['scores'], rtol=2e-6, atol=2e-7, err_msg='step %d' % t)
        pos = list(tr['candidates']).index(ret[t])
        assert tr['scores'][pos] >= tr['scores'].max() - 1e-6 * abs(tr['scores'].max())
    assert gpu.fetch_unlabelled(3) == ret
    with pytest.raises(NotImplementedError):
        gpu.fetch_unlabelled(6)


def test_optimistic_golden_from_the_reference():
    """The reference's own record for label_estimation='optimistic' (label_prob = 1, mistake_prob = 0.2); loose like
    the oracle's comparison (tests/test_oracle_golden.py: the winning term is set by the relative accuracy of small
    probabilities, ill-conditioned in the reference itself)."""
    g = load_golden('butterflies_optimistic_k2')
    kw = dict(g['learner_kw'])
    gpu = drive(_gpu_learner(g['X'], **kw), g)
    ret = gpu._fetch_stepwise(int(g['k']), keep_scores=True)
    for t, (sc, st) in enumerate(zip(gpu.last_step_scores, g['steps'])):
        np.testing.assert_allclose(sc[st['candidates']], st['mi'], rtol=1e-2, atol=1e-5, err_msg='step %d' % t)
        if ret[t] != st['chosen']:
            break


def _subset_problem(n=60, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.randn(n, 2)
    y = np.where(X[:, 0] + 0.3 * X[:, 1] > 0, 1, -1)
    return X, {0: int(y[0]), 5: int(y[5]), 9: int(y[9]), 14: int(y[14])}


@pytest.mark.parametrize('ce,k,noise,mp', [(3, 3, 1e-6, 0.0), (5, 4, 1e-6, 0.0), (2, 4, 1e-3, 0.0), (4, 2, 1e-6, 0.0),
                                           (3, 3, 1e-6, 0.25)])
def test_change_estimation_subset_matches_the_oracle(ce, k, noise, mp):
    """ITAL(change_estimation_subset = c) (ital.py:102-108, 227-275): the subset comes from the reference's own draw
    on the global numpy RNG; every candidate's score of every step within 1e-6 of the oracle's restatement of the
    kernel's form (oracle/ce_subset.py mi_sub_shared; up to six variables -- beyond that the lattice of the prior
    differs in its 1e-15 inverse-CDF round-off only), same batch."""
    from oracle.ital_oracle import OracleITAL
    X, fb = _subset_problem()
    kw = dict(length_scale=1.0, noise=noise, change_estimation_subset=ce, mistake_prob=mp)
    gpu, ora = _gpu_learner(X, **kw), OracleITAL(X, **kw)
    gpu.update(fb)
    ora.update(fb)
    np.random.seed(123)
    state = np.random.get_state()
    ret = gpu._fetch_change_subset(k, keep_scores=True)
    np.random.set_state(state)
    want = ora.fetch_unlabelled(k, forced=ret)
    assert gpu.last_subset == ora.subset and want == ret
    for t, tr in enumerate(ora.trace):
        got = gpu.last_step_scores[t][tr['candidates']]
        assert not np.any(np.isnan(got)), 'step %d: unscored candidates' % t
        np.testing.assert_allclose(got, tr['scores'], rtol=2e-6, atol=1e-9, err_msg='step %d' % t)
        best = float(np.max(tr['scores']))
        assert tr['scores'][list(tr['candidates']).index(ret[t])] >= best - 1e-9 * max(1.0, abs(best))
    # the public call draws the same subset and returns the same batch
    np.random.set_state(state)
    assert gpu.fetch_unlabelled(k) == ret
    gpu.close()


@pytest.mark.parametrize('name', subset_names())
def test_change_estimation_subset_matches_the_reference(name):
    """The goldens of the unmodified reference with change_estimation_subset (make_subset_golden.py): same subset,
    same batch, every candidate's score of every step within 1e-4 (see test_oracle_golden.py for the tolerance)."""
    g = load_subset(name)
    gpu = _gpu_learner(g['X'], length_scale=float(g['length_scale']), var=float(g['var']), noise=float(g['noise']),
                       change_estimation_subset=int(g['change_estimation_subset']), mistake_prob=float(g['mistake_prob']))
    for fb in g['updates']:
        gpu.update({int(k): v for k, v in fb.items()})
    np.random.seed(int(g['seed']))
    ret = gpu._fetch_change_subset(int(g['k']), keep_scores=True)
    assert gpu.last_subset == [int(i) for i in g['subset']]
    assert ret == [int(i) for i in g['ret']]
    for t, st in enumerate(g['steps']):
        got = gpu.last_step_scores[t][st['candidates']]
        six = len(g['subset']) + t + 1 >= 6              # (five base variables at Q = 10: see test_oracle_golden.py)
        loose = float(g['mistake_prob']) > 0             # (one near-duplicate of a subset member: see test_oracle_golden.py)
        assert np.sum(np.abs(got - st['mi']) > 1e-4 + 1e-4 * np.abs(st['mi'])) <= 2
        np.testing.assert_allclose(got, st['mi'], rtol=1e-4, atol=1.0 if loose else (1e-3 if six else 1e-4),
                                   err_msg='step %d' % t)
    gpu.close()


@pytest.mark.parametrize('name', clip_names())
def test_clip_cov_matches_the_reference(name):
    """ITAL(clip_cov = th) with 6 samples per batch against the goldens of the unmodified reference: same batch, the
    scores of the grouped step (more than 5 samples, ital.py:360-362) within 1e-6."""
    g = load_clip(name)
    gpu = _gpu_learner(g['X'], length_scale=float(g['length_scale']), var=float(g['var']), noise=float(g['noise']),
                       clip_cov=float(g['clip_cov']))
    for fb in g['updates']:
        gpu.update({int(k): v for k, v in fb.items()})
    assert gpu.fetch_unlabelled(int(g['k'])) == [int(i) for i in g['ret']]
    ret = gpu._fetch_stepwise(int(g['k']), keep_scores=True)
    assert ret == [int(i) for i in g['ret']]
    st = g['steps'][5]
    got = gpu.last_step_scores[5][st['candidates']]
    np.testing.assert_allclose(got, st['mi'], rtol=1e-6, atol=1e-6)
    gpu.close()


@pytest.mark.parametrize('th,ls,k', [(0.3, 1.0, 8), (0.15, 0.7, 7), (0.6, 1.5, 6)])
def test_clip_cov_tracks_the_oracle(th, ls, k):
    """Longer batches with clip_cov: every candidate's score of every grouped step within 1e-6 of the oracle's literal
    product over the groups (oracle/ital_oracle.py mi_grouped), same batch."""
    from oracle.ital_oracle import OracleITAL
    rng = np.random.RandomState(11)
    X = rng.randn(80, 2) * 1.5
    y = np.where(X[:, 0] - 0.3 * X[:, 1] > 0, 1, -1)
    fb = {i: int(y[i]) for i in (0, 7, 19, 33, 50)}
    gpu, ora = _gpu_learner(X, length_scale=ls, clip_cov=th), OracleITAL(X, length_scale=ls, clip_cov=th)
    gpu.update(fb)
    ora.update(fb)
    ret = gpu._fetch_stepwise(k, keep_scores=True)
    ora.fetch_unlabelled(k, forced=ret)
    for t in range(5, k):
        tr = ora.trace[t]
        got = gpu.last_step_scores[t][tr['candidates']]
        np.testing.assert_allclose(got, tr['scores'], rtol=1e-6, atol=1e-9, err_msg='step %d' % t)
        best = float(np.max(tr['scores']))
        assert tr['scores'][list(tr['candidates']).index(ret[t])] >= best - 1e-9 * max(1.0, abs(best))
    assert gpu.fetch_unlabelled(k) == ret
    gpu.close()


def test_clip_cov_with_a_user_who_mislabels():
    """label_prob = 1, mistake_prob > 0 with clip_cov: the perfect-user scores plus the per-step constant
    (1 - (1 - mp)^D) (log eps - log(1 + eps)), so the same batch."""
    rng = np.random.RandomState(11)
    X = rng.randn(80, 2) * 1.5
    y = np.where(X[:, 0] - 0.3 * X[:, 1] > 0, 1, -1)
    fb = {i: int(y[i]) for i in (0, 7, 19, 33, 50)}
    a, b = _gpu_learner(X, length_scale=1.0, clip_cov=0.3), _gpu_learner(X, length_scale=1.0, clip_cov=0.3, mistake_prob=0.2)
    a.update(fb)
    b.update(fb)
    ra, rb = a.fetch_unlabelled(7), b.fetch_unlabelled(7)
    assert ra == rb
    eps = 1e-12
    shift = [(1 - 0.8 ** (t + 1)) * (np.log(eps) - np.log1p(eps)) for t in range(7)]
    np.testing.assert_allclose(np.array(b.last_fetch_scores) - np.array(a.last_fetch_scores), shift, rtol=1e-6)
    a.close()
    b.close()
