"""Goldens of the two baseline learners that share ITAL's kernels, recorded from the
UNMODIFIED reference: EntropySampling and VarianceSampling (/root/reference/ital/baseline_methods.py:110-155, 229-287).

    python tests/golden/make_baseline_golden.py

For EntropySampling the greedy loop of fetch_unlabelled (baseline_methods.py:241-261) is replayed around the
reference's own single_entropy / batch_entropy so that the per-candidate entropies of every step are kept, and the
result is checked against the reference's fetch_unlabelled itself.  Outputs tests/golden/baseline_<case>.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

ital = ref_shims.install()
from ital.baseline_methods import EntropySampling, VarianceSampling  # noqa: E402

sys.path.insert(0, HERE)
from make_golden import butterflies, labelled_rounds, syn_pool  # noqa: E402


def entropy_steps(learner, k):
    """baseline_methods.py:241-261 with the per-step entropies kept."""
    rel_mean, rel_var = learner.gp.predict_stored(cov_mode='diag')
    rel_mean, rel_var = rel_mean[:len(learner.data)], rel_var[:len(learner.data)]
    candidates = learner.get_unseen()
    ent0 = np.array([EntropySampling.single_entropy(rel_mean[i], rel_var[i]) for i in candidates])
    max_ind = int(np.argmax(ent0))
    steps = [dict(candidates=np.array(candidates), entropy=ent0, chosen=candidates[max_ind])]
    ret = [candidates[max_ind]]
    for l in range(1, k):
        del candidates[max_ind]
        covs = learner.gp.predict_cov_batch(ret, candidates)
        ent = np.array([EntropySampling.batch_entropy(rel_mean[ret + [candidates[i]]], covs[i]) for i in range(len(candidates))])
        max_ind = int(np.argmax(ent))
        steps.append(dict(candidates=np.array(candidates), entropy=ent, chosen=candidates[max_ind]))
        ret.append(candidates[max_ind])
    return ret, steps


def run(name, X, updates, k, kw, unnameable=()):
    out = dict(X=X, k=k, n_updates=len(updates), unnameable=np.array(list(unnameable), dtype=np.int64))
    for key in ('length_scale', 'var', 'noise'):
        out[key] = float(kw.get(key, dict(length_scale=0.1, var=1.0, noise=1e-6)[key]))
    for u, fb in enumerate(updates):
        out['upd%d_idx' % u] = np.array(list(fb.keys()), dtype=np.int64)
        out['upd%d_val' % u] = np.array(list(fb.values()), dtype=np.float64)

    def prepared(cls, **extra):
        L = cls(X, **kw, **extra)
        for fb in updates:
            L.update(fb)
        if len(unnameable):
            L.update({int(i): 0 for i in unnameable})
        return L

    ent = prepared(EntropySampling)
    ret, steps = entropy_steps(ent, k)
    assert ret == [int(i) for i in prepared(EntropySampling).fetch_unlabelled(k)], 'replay differs from fetch_unlabelled'
    out['entropy_ret'] = np.array(ret, dtype=np.int64)
    for t, st in enumerate(steps):
        out['entropy_step%d_candidates' % t] = st['candidates']
        out['entropy_step%d_entropy' % t] = st['entropy']
    out['variance_ret'] = np.array(prepared(VarianceSampling).fetch_unlabelled(k), dtype=np.int64)
    out['variance_corr_ret'] = np.array(prepared(VarianceSampling, use_correlations=True).fetch_unlabelled(k), dtype=np.int64)
    out['var_diag'] = ent.gp.predict_stored(cov_mode='diag')[1][:len(X)]
    np.savez_compressed(os.path.join(HERE, 'baseline_%s.npz' % name), **out)
    print('%-22s n=%d k=%d entropy %s variance %s variance(corr) %s' % (
        name, len(X), k, ret, out['variance_ret'].tolist(), out['variance_corr_ret'].tolist()), flush=True)


if __name__ == '__main__':
    rng = np.random.default_rng(20181011)
    Xb, yb = butterflies()
    sub = np.sort(rng.choice(len(Xb), 200, replace=False))
    run('butterflies_sub', Xb[sub], labelled_rounds(yb[sub], int(yb[sub][0]), rng, rounds=2), 4, dict(length_scale=2.5),
        unnameable=[7, 19])
    Xs, assign = syn_pool(600, d=64, centres=12)
    ys = (assign == assign[0]).astype(int)
    upd = [{0: 1}, {int(i): (1 if ys[i] else -1) for i in list(np.nonzero(ys)[0][1:4]) + list(np.nonzero(1 - ys)[0][:4])}]
    run('syn600', Xs, upd, 4, dict(length_scale=1.0))
