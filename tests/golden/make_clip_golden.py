"""Golden vectors of ITAL(clip_cov = th) with batches of more than 5 samples, from the UNMODIFIED reference.

    python tests/golden/make_clip_golden.py

Same arrangement as make_golden.py (reference imported read-only through oracle/ref_shims).  From the sixth sample of a
batch on, MutualInformation.prob_rel factorises the orthant probability over the groups of samples whose correlations
exceed clip_cov (ital/ital.py:360-362, 386-429, 590-616).  The greedy loop is replayed around the reference's own
AppendedMutualInformation (make_golden.replay_fetch) so that every candidate's score of every step is kept.

Outputs: tests/golden/clip_<case>.npz.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shims, imports the reference)


def run_case(name, X, updates, k, kw):
    t0 = time.time()
    learner = mg.ital.ITAL(X, parallelized=False, **kw)
    for fb in updates:
        learner.update(fb)
    ret, steps = mg.replay_fetch(learner, k)
    out = dict(X=X, k=k, ret=np.array(ret, dtype=np.int64), rel_mean=np.array(learner.rel_mean),
               n_updates=len(updates), clip_cov=float(kw['clip_cov']))
    for key in ('length_scale', 'var', 'noise'):
        out[key] = float(kw.get(key, dict(length_scale=0.1, var=1.0, noise=1e-6)[key]))
    for u, fb in enumerate(updates):
        out['upd%d_idx' % u] = np.array(list(fb.keys()), dtype=np.int64)
        out['upd%d_val' % u] = np.array(list(fb.values()), dtype=np.float64)
    for t, s in enumerate(steps):
        for key in ('candidates', 'mi', 'chosen'):
            out['step%d_%s' % (t, key)] = s[key]
    np.savez_compressed(os.path.join(HERE, 'clip_' + name + '.npz'), **out)
    print('%-20s n=%d k=%d ret=%s  %.1fs' % (name, X.shape[0], k, ret, time.time() - t0), flush=True)


if __name__ == '__main__':
    rng = np.random.RandomState(3)
    X = rng.randn(36, 2) * 1.5
    y = np.where(X[:, 0] + 0.5 * X[:, 1] > 0, 1, -1)
    upd = [{0: int(y[0]), 5: int(y[5]), 9: int(y[9]), 20: int(y[20])}]
    run_case('randn_th03_k6', X, upd, 6, dict(length_scale=1.0, noise=1e-6, clip_cov=0.3))
    run_case('randn_th01_k6', X, upd, 6, dict(length_scale=0.6, noise=1e-6, clip_cov=0.1))
