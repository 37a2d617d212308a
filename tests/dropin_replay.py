"""Replay of a recorded run of the reference's experiment driver (tests/golden/make_experiment_golden.py) on another
learner with the reference's interface -- shared by the CPU test (oracle) and the GPU test (ital_b200.ITAL)."""
import json
import math
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_experiment(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'experiment_%s.npz' % name), allow_pickle=False))
    g['log'] = json.loads(str(g['log']))
    g['learner_kw'] = json.loads(str(g['learner_kw']))
    g['classes'] = json.loads(str(g['classes']))
    g['table'] = str(g['table']).splitlines()
    return g


def ndcg(y_true, y_score):
    """utils.ndcg of the reference (utils.py:174-200), restated."""
    num_relevant = sum(yt > 0 for yt in y_true)
    rank, cgain, normalizer = 0, 0.0, 0.0
    for ret in np.argsort(y_score)[::-1]:
        if y_true[ret] != 0:
            rank += 1
            gain = 1.0 / math.log2(rank + 1)
            if y_true[ret] > 0:
                cgain += gain
            if rank <= num_relevant:
                normalizer += gain
    return cgain / normalizer


def replay(learner, g, checker=None, predict_rtol=1e-6, predict_atol=1e-8):
    """Drive `learner` through the recorded calls (run_experiment.py:145-168).  The batches must be identical, or --
    where the reference's own maximum is a floating-point tie (toy data: candidates far from every labelled point
    have the prior's score to 1e-15) -- the recorded batch must be a maximiser too: `checker`, an oracle learner kept
    in the same state, re-scores the recorded greedy path and every recorded choice has to reach its maximum within
    1e-9 relative.  The recorded feedback is applied either way, so one tie does not derail the rest of the run.  The
    test-set predictions must agree within tolerance.  Returns per query the AP / NDCG lists computed from THIS
    learner's predictions, in the order the driver went through classes and queries, and the number of tie breaks."""
    from sklearn.metrics import average_precision_score
    log = g['log']
    cls = iter(g['classes'])
    aps, ndcgs = [], []
    rel_test = None
    n_resets = sum(1 for e in log if e['op'] == 'reset')
    per_class = n_resets // len(g['classes'])
    seen_resets = 0
    ties = 0
    for e in log:
        if e['op'] == 'reset':
            if seen_resets % per_class == 0:
                rel_test = np.asarray(g['rel_test_' + next(cls)])
            seen_resets += 1
            learner.reset()
            if checker is not None:
                checker.reset()
            aps.append([])
            ndcgs.append([])
        elif e['op'] == 'update':
            fb = {i: v for i, v in zip(e['idx'], e['val'])}
            learner.update(fb)
            if checker is not None:
                checker.update(fb)
        elif e['op'] == 'fetch':
            ret = [int(i) for i in learner.fetch_unlabelled(e['k'])]
            assert len(ret) == len(e['ret']) == len(set(ret))
            if ret != e['ret']:
                assert checker is not None, (e, ret)
                checker.fetch_unlabelled(e['k'], forced=e['ret'])
                for t, tr in enumerate(checker.trace):
                    pos = int(np.nonzero(tr['candidates'] == e['ret'][t])[0][0])
                    assert tr['scores'][pos] >= tr['scores'].max() - 1e-9 * abs(tr['scores'].max()), (e, ret, t)
                ties += 1
        elif e['op'] == 'predict':
            scores = np.asarray(learner.gp.predict(g['X_test']))
            np.testing.assert_allclose(scores, g[e['out']], rtol=predict_rtol, atol=predict_atol)
            nz = rel_test != 0
            aps[-1].append(average_precision_score(rel_test[nz], scores[nz]))
            ndcgs[-1].append(ndcg(rel_test, scores))
    return np.array(aps), np.array(ndcgs), ties


def table_of(aps, ndcgs):
    """The `Round;Median_AP;...` lines of run_experiment.py:193-195 (avg_class_perf = yes: all queries pooled)."""
    out = ['Round;Median_AP;Mean_AP;AP_SD;Median_NDCG;Mean_NDCG;NDCG_SD']
    for i in range(aps.shape[1]):
        out.append('{};{:.4f};{:.4f};{:.4f};{:.4f};{:.4f};{:.4f}'.format(
            i, np.median(aps[:, i]), np.mean(aps[:, i]), np.std(aps[:, i]), np.median(ndcgs[:, i]), np.mean(ndcgs[:, i]),
            np.std(ndcgs[:, i])))
    return out
