import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    g['updates'] = [dict(zip(g['upd%d_idx' % u].tolist(), g['upd%d_val' % u].tolist()))
                    for u in range(int(g['n_updates']))]
    g['steps'] = []
    t = 0
    while 'step%d_mi' % t in g:
        g['steps'].append(dict(candidates=g['step%d_candidates' % t], mi=g['step%d_mi' % t],
                               chosen=int(g['step%d_chosen' % t]), rel_covs=g['step%d_rel_covs' % t]))
        t += 1
    tc = float(g['top_candidates'])
    g['learner_kw'] = dict(length_scale=float(g['length_scale']), var=float(g['var']), noise=float(g['noise']),
                           label_prob=float(g['label_prob']), mistake_prob=float(g['mistake_prob']),
                           label_estimation=str(g['label_estimation']),
                           top_candidates=None if tc < 0 else (int(tc) if tc == int(tc) else tc))
    return g


def drive(learner, g):
    """Bring a learner (oracle or product) to the labelled state the golden was recorded in."""
    for fb in g['updates']:
        learner.update(fb)
    if len(g['unnameable']):
        learner.update({int(i): 0 for i in g['unnameable']})
    return learner


@pytest.fixture(params=golden_names())
def golden(request):
    return request.param, load_golden(request.param)
