"""Golden vectors of ITAL(change_estimation_subset = c) from the UNMODIFIED reference (/root/reference).

    python tests/golden/make_subset_golden.py

Same arrangement as make_golden.py (reference imported read-only through oracle/ref_shims, deterministic stand-in for
the removed mvndst): the subset is drawn by the reference's own line (ital/ital.py:105-106) from the global numpy RNG
seeded here, the greedy loop of ITAL.fetch_unlabelled (ital.py:119-132) is replayed around the reference's own
AppendedMutualInformation so that every candidate's score of every step is kept, and the replay is checked against
``fetch_unlabelled`` itself under the same seed.  Cases keep batch + subset at six variables or fewer: that is what the
stand-in's tensor rule covers (the reference's own mvndst(maxpts = 100 dim) is 1e-3 noise there).

Outputs: tests/golden/subset_<case>.npz.
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shims, imports the reference)

from ital.ital import AppendedMutualInformation, _init_pool, _parallel_mi  # noqa: E402


def replay(learner, k, seed, procs=8):
    np.random.seed(seed)
    candidates = learner.get_unseen()
    k = min(k, len(candidates))
    learner._ce_subset = sorted(np.random.choice(candidates, min(len(candidates), learner.change_estimation_subset),
                                                 replace=False))                   # ital.py:105-106
    mutual_information = AppendedMutualInformation(learner)
    steps = []
    for it in range(k):
        with Pool(procs, initializer=_init_pool, initargs=(mutual_information,)) as p:
            mi = p.map(_parallel_mi, candidates)
        max_ind = int(np.argmax(mi))
        steps.append(dict(candidates=np.array(candidates, dtype=np.int64), mi=np.array(mi, dtype=np.float64),
                          chosen=int(candidates[max_ind])))
        mutual_information.append(candidates[max_ind])
        del candidates[max_ind]
    return mutual_information.ret, [int(i) for i in learner._ce_subset], steps


def run_case(name, X, updates, k, seed, kw):
    t0 = time.time()
    learner = mg.ital.ITAL(X, parallelized=False, **kw)
    for fb in updates:
        learner.update(fb)
    ret, subset, steps = replay(learner, k, seed)
    np.random.seed(seed)
    assert learner.fetch_unlabelled(k) == ret, 'replay differs from ITAL.fetch_unlabelled'
    out = dict(X=X, k=k, seed=seed, ret=np.array(ret, dtype=np.int64), subset=np.array(subset, dtype=np.int64),
               rel_mean=np.array(learner.rel_mean), n_updates=len(updates),
               change_estimation_subset=int(kw['change_estimation_subset']),
               mistake_prob=float(kw.get('mistake_prob', 0.0)))
    for key in ('length_scale', 'var', 'noise'):
        out[key] = float(kw.get(key, dict(length_scale=0.1, var=1.0, noise=1e-6)[key]))
    for u, fb in enumerate(updates):
        out['upd%d_idx' % u] = np.array(list(fb.keys()), dtype=np.int64)
        out['upd%d_val' % u] = np.array(list(fb.values()), dtype=np.float64)
    for t, s in enumerate(steps):
        for key, v in s.items():
            out['step%d_%s' % (t, key)] = v
    np.savez_compressed(os.path.join(HERE, 'subset_' + name + '.npz'), **out)
    print('%-24s n=%d k=%d subset=%s ret=%s  %.1fs' % (name, X.shape[0], k, subset, ret, time.time() - t0), flush=True)


if __name__ == '__main__':
    want = set(sys.argv[1:])
    rng = np.random.default_rng(20181010)
    Xt, yt = mg.toy_data()
    pick = np.sort(rng.choice(len(Xt), 48, replace=False))
    X, y = Xt[pick], yt[pick]
    upd = mg.labelled_rounds(y, 1, rng, rounds=1)
    Xb, yb = mg.butterflies()
    sub = np.sort(rng.choice(len(Xb), 40, replace=False))
    upd_b = mg.labelled_rounds(yb[sub], int(yb[sub][0]), rng, 1)
    for name, args in (('toy_c3_k3', (X, upd, 3, 11, dict(length_scale=0.1, change_estimation_subset=3))),
                       ('toy_c2_k4', (X, upd, 4, 5, dict(length_scale=0.1, change_estimation_subset=2))),
                       ('butterflies_c3_k2', (Xb[sub], upd_b, 2, 3, dict(length_scale=2.5, change_estimation_subset=3))),
                       ('toy_c2_k3_mp02', (X, upd, 3, 9, dict(length_scale=0.1, change_estimation_subset=2,
                                                              label_prob=1.0, mistake_prob=0.2)))):
        if not want or name in want:
            run_case(name, *args)
