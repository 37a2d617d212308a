"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in this container.

    python tests/golden/make_golden.py [case ...]

The reference has no tests or fixtures of its own, so the goldens are outputs of the
reference code itself: ``ITAL`` / ``AppendedMutualInformation`` / ``GaussianProcess`` imported read-only
from /root/reference through ``oracle/ref_shims`` (numpy ``numexpr`` stub; deterministic high-order stand-in
for the removed ``scipy.stats.mvn.mvndst``).  The greedy loop of ``ITAL.fetch_unlabelled``
(ital/ital.py:119-132) is replayed line by line around the reference's own scoring object so that the full
per-candidate MI vector of every greedy step can be frozen, not only the chosen indices; the test
``test_golden_fetch_matches_reference_fetch`` checks that replay against ``fetch_unlabelled`` itself.

Outputs: tests/golden/<case>.npz.  /root/reference does not exist on the GPU box; only these files travel.
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

ital = ref_shims.install()
from ital.ital import AppendedMutualInformation, _init_pool, _parallel_mi  # noqa: E402


def syn_pool(n, d=512, centres=1000, seed=0):
    """SYN pool of SURVEY.md section 8(d): clustered, L2-normalised, float32-representable."""
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((centres, d))
    assign = rng.integers(0, centres, n)
    X = C[assign] + 0.6 * rng.standard_normal((n, d))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    return X.astype(np.float32).astype(np.float64), assign


def toy_data():
    sys.path.insert(1, ref_shims.REFERENCE_ROOT)
    import datasets
    ds = datasets.ToyDataset(size_factor=10)          # configs/toy.conf:10-11 -> 170 points, 85 train
    return ds.X_train_norm, ds.y_train


def butterflies():
    d = np.load(os.path.join(ref_shims.REFERENCE_ROOT, 'data', 'butterflies_pca50.npz'))
    X = d['X_train']
    return (X - X.min()) / (X.max() - X.min()), d['y_train']      # datasets.py:107-112


def usps_test():
    import datasets
    ds = datasets.USPSDataset.__new__(datasets.USPSDataset)
    X, y = ds._read_usps(os.path.join(ref_shims.REFERENCE_ROOT, 'data', 'usps_test.jf'))
    return (X - X.min()) / (X.max() - X.min()), y


def replay_fetch(learner, k, procs=8):
    """ital/ital.py:98-134 with the per-step MI vectors kept."""
    candidates = learner.get_unseen()
    k = min(k, len(candidates))
    learner._ce_subset = None
    if learner.top_candidates is not None:
        top = learner.top_candidates
        if isinstance(top, float):
            top = min(len(candidates), int(top * (len(learner.queries) + len(learner.relevant_ids)
                                                  + len(learner.irrelevant_ids))))
        if (top > 0) and (top < len(candidates)):
            top_ind = np.argpartition(learner.rel_mean[candidates], -top)[-top:]
            candidates = [candidates[i] for i in top_ind]
    mutual_information = AppendedMutualInformation(learner)
    steps = []
    for it in range(k):
        with Pool(procs, initializer=_init_pool, initargs=(mutual_information,)) as p:
            mi = p.map(_parallel_mi, candidates)
        max_ind = int(np.argmax(mi))
        steps.append(dict(candidates=np.array(candidates, dtype=np.int64), mi=np.array(mi, dtype=np.float64),
                          chosen=int(candidates[max_ind]),
                          rel_covs=np.array(mutual_information.rel_covs[candidates], dtype=np.float64)))
        mutual_information.append(candidates[max_ind])
        del candidates[max_ind]
    return mutual_information.ret, steps


def run_case(name, X, updates, k, learner_kw, unnameable=(), queries=()):
    t0 = time.time()
    kw = dict(learner_kw)
    learner = ital.ITAL(X, queries=list(queries), parallelized=False, **kw)
    for fb in updates:
        learner.update(fb)
    if len(unnameable):
        learner.update({int(i): 0 for i in unnameable})
    ret, steps = replay_fetch(learner, k)
    out = dict(X=X, k=k, ret=np.array(ret, dtype=np.int64),
               rel_mean=np.array(learner.rel_mean), gp_ind=np.array(learner.gp.ind, dtype=np.int64),
               gp_y=np.array(learner.gp.y, dtype=np.float64), gp_w=np.array(learner.gp.w),
               var_diag=learner.gp.predict_stored(cov_mode='diag')[1][:len(X)],
               unnameable=np.array(list(unnameable), dtype=np.int64),
               queries=np.array(queries, dtype=np.float64).reshape(len(queries), X.shape[1]),
               n_updates=len(updates))
    for key in ('length_scale', 'var', 'noise', 'label_prob', 'mistake_prob'):
        out[key] = float(kw.get(key, dict(length_scale=0.1, var=1.0, noise=1e-6, label_prob=1.0,
                                          mistake_prob=0.0)[key]))
    out['label_estimation'] = str(kw.get('label_estimation', 'mean'))
    out['top_candidates'] = -1 if kw.get('top_candidates') is None else kw['top_candidates']
    for u, fb in enumerate(updates):
        out['upd%d_idx' % u] = np.array(list(fb.keys()), dtype=np.int64)
        out['upd%d_val' % u] = np.array(list(fb.values()), dtype=np.float64)
    for t, s in enumerate(steps):
        for key, v in s.items():
            out['step%d_%s' % (t, key)] = v
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('%-28s n=%d d=%d k=%d ret=%s  %.1fs' % (name, X.shape[0], X.shape[1], k, ret, time.time() - t0),
          flush=True)


def run_updated_prediction(name, X, updates, learner_kw, probes, queries=()):
    """ActiveRetrievalBase.updated_prediction (retrieval_base.py:129-164 -> gp.py:295-344) of the unmodified
    reference for a few (feedback, test_ind) probes, all three cov_modes."""
    learner = ital.ITAL(X, queries=list(queries), parallelized=False, **learner_kw)
    for fb in updates:
        learner.update(fb)
    out = dict(X=X, n_updates=len(updates), n_probes=len(probes),
               queries=np.array(queries, dtype=np.float64).reshape(len(queries), X.shape[1]))
    for key in ('length_scale', 'var', 'noise'):
        out[key] = float(learner_kw.get(key, dict(length_scale=0.1, var=1.0, noise=1e-6)[key]))
    for u, fb in enumerate(updates):
        out['upd%d_idx' % u] = np.array(list(fb.keys()), dtype=np.int64)
        out['upd%d_val' % u] = np.array(list(fb.values()), dtype=np.float64)
    for p, (fb, test_ind) in enumerate(probes):
        out['probe%d_fb_idx' % p] = np.array(list(fb.keys()), dtype=np.int64)
        out['probe%d_fb_val' % p] = np.array(list(fb.values()), dtype=np.float64)
        out['probe%d_test' % p] = np.array(test_ind, dtype=np.int64)
        out['probe%d_mean' % p] = np.array(learner.updated_prediction(fb, test_ind, cov_mode=None))
        m, v = learner.updated_prediction(fb, test_ind, cov_mode='diag')
        out['probe%d_var' % p] = np.array(v)
        m, c = learner.updated_prediction(fb, test_ind, cov_mode='full')
        out['probe%d_cov' % p] = np.array(c)
        assert np.allclose(m, out['probe%d_mean' % p])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('%-28s n=%d probes=%d' % (name, X.shape[0], len(probes)), flush=True)


def labelled_rounds(y, positive, rng, rounds=2, per_round=4):
    """A query plus a few feedback rounds, like run_experiment.py:142-164 (simulated perfect feedback)."""
    pos = np.nonzero(y == positive)[0]
    q = int(rng.choice(pos))
    updates = [{q: 1}]
    seen = {q}
    for _ in range(rounds):
        fb = {}
        while len(fb) < per_round:
            i = int(rng.integers(0, len(y)))
            if i not in seen:
                seen.add(i)
                fb[i] = 1 if y[i] == positive else -1
        updates.append(fb)
    return updates


def cases():
    rng = np.random.default_rng(20181009)
    Xt, yt = toy_data()
    upd_t = labelled_rounds(yt, 1, rng, rounds=1)
    yield 'toy_perfect_k4', (Xt, upd_t, 4, dict(length_scale=0.1)), {}
    yield 'toy_mistakes_k3', (Xt, upd_t, 3, dict(length_scale=0.1, label_prob=0.75, mistake_prob=0.2)), {}
    yield 'toy_queries_k3', (Xt[:60], [], 3, dict(length_scale=0.1)), dict(queries=[Xt[70], Xt[80]])
    yield 'toy_topcand_k3', (Xt, upd_t, 3, dict(length_scale=0.1, top_candidates=20)), {}

    Xb, yb = butterflies()
    upd_b = labelled_rounds(yb, int(yb[0]), rng, rounds=2)
    yield 'butterflies_k2', (Xb, upd_b, 2, dict(length_scale=2.5)), dict(unnameable=[5, 17])
    sub = np.sort(rng.choice(len(Xb), 160, replace=False))
    upd_bs = labelled_rounds(yb[sub], int(yb[sub][0]), rng, rounds=2)
    yield 'butterflies_sub_k4', (Xb[sub], upd_bs, 4, dict(length_scale=2.5)), {}
    yield 'butterflies_aggressive_k3', (Xb[sub][:90], labelled_rounds(yb[sub][:90], int(yb[sub][0]), rng, 1), 3,
                                        dict(length_scale=2.5, label_prob=1.0, mistake_prob=0.5)), {}
    yield 'butterflies_conservative_k3', (Xb[sub][:90], labelled_rounds(yb[sub][:90], int(yb[sub][0]), rng, 1),
                                          3, dict(length_scale=2.5, label_prob=0.25, mistake_prob=0.0)), {}
    yield 'butterflies_optimistic_k2', (Xb[sub][:90], labelled_rounds(yb[sub][:90], int(yb[sub][0]), rng, 1), 2,
                                        dict(length_scale=2.5, label_prob=1.0, mistake_prob=0.2,
                                             label_estimation='optimistic')), {}

    Xu, yu = usps_test()
    subu = np.sort(rng.choice(len(Xu), 400, replace=False))
    yield 'usps_sub_k2', (Xu[subu], labelled_rounds(yu[subu], int(yu[subu][0]), rng, 2), 2,
                          dict(length_scale=3.0)), {}
    yield 'usps_sub_k4', (Xu[subu][:150], labelled_rounds(yu[subu][:150], int(yu[subu][0]), rng, 2), 4,
                          dict(length_scale=3.0)), {}

    Xs, assign = syn_pool(2000, d=64, centres=40)
    ys = (assign == assign[0]).astype(int)
    upd_s = [{0: 1}, {int(i): (1 if ys[i] else -1) for i in
                      list(np.nonzero(ys)[0][1:5]) + list(np.nonzero(1 - ys)[0][:4])}]
    yield 'syn2000_k2', (Xs, upd_s, 2, dict(length_scale=1.0)), {}
    yield 'syn200_k4', (Xs[:200].copy(), [{0: 1}, {3: -1, 9: -1, 11: 1 if ys[11] else -1, 20: -1}], 4,
                        dict(length_scale=1.0)), {}
    # configs/toy.conf ships batch_size = 6: the fifth and sixth pick use five and six variables (tensor rule on the
    # device for 4 and 5 base variables); the stand-in takes ~1 s per six-variable probability, ~15 min for this case
    yield 'toy_perfect_k6', (Xt, upd_t, 6, dict(length_scale=0.1)), {}


def updated_prediction_cases():
    rng = np.random.default_rng(20181010)
    Xt, yt = toy_data()
    upd_t = labelled_rounds(yt, 1, rng, rounds=1)
    seen = set(i for fb in upd_t for i in fb)
    free = [i for i in range(len(Xt)) if i not in seen]
    probes = [({free[0]: 1}, [free[0], free[1]]),
              ({free[3]: -1, free[2]: 1, free[5]: 0}, [free[2], free[3], free[5], free[7]]),
              ({free[4]: 0}, [free[4], free[6], free[8]]),                       # nothing annotated: predict_stored
              ({free[10]: 1, free[9]: 1, free[12]: -1, free[11]: -1}, [free[9], free[10], free[11], free[12]])]
    yield 'updpred_toy', (Xt, upd_t, dict(length_scale=0.1), probes), {}
    Xs, assign = syn_pool(300, d=64, centres=10)
    upd_s = [{0: 1}, {3: -1, 9: -1, 11: 1 if assign[11] == assign[0] else -1, 20: -1}]
    probes = [({50: 1, 40: -1}, [40, 50, 60, 299]), ({100: -1}, [100]), ({7: 1, 8: 1, 12: -1}, [7, 8, 12, 13, 14])]
    yield 'updpred_syn300', (Xs, upd_s, dict(length_scale=1.0), probes), dict(queries=[Xs[150] + 0.01])


if __name__ == '__main__':
    want = set(sys.argv[1:])
    for name, args, kw in updated_prediction_cases():
        if want and name not in want:
            continue
        run_updated_prediction(name, *args, **kw)
    if want and all(w.startswith('updpred') for w in want):
        sys.exit(0)
    for name, args, kw in cases():
        if want and name not in want:
            continue
        run_case(name, *args, **kw)
