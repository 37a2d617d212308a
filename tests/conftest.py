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


def _npz_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def golden_names():
    """Fetch goldens (every greedy step of a fetch_unlabelled of the reference)."""
    return [n for n in _npz_names() if not n.startswith(('updpred_', 'experiment_', 'baseline_', 'subset_', 'clip_'))]


def updpred_names():
    """updated_prediction goldens (make_golden.py run_updated_prediction)."""
    return [n for n in _npz_names() if n.startswith('updpred_')]


def subset_names():
    """Goldens of ITAL(change_estimation_subset = c) (make_subset_golden.py)."""
    return [n for n in _npz_names() if n.startswith('subset_')]


def load_subset(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    g['updates'] = [dict(zip(g['upd%d_idx' % u].tolist(), g['upd%d_val' % u].tolist()))
                    for u in range(int(g['n_updates']))]
    g['steps'] = [dict(candidates=g['step%d_candidates' % t], mi=g['step%d_mi' % t], chosen=int(g['step%d_chosen' % t]))
                  for t in range(len(g['ret']))]
    g.setdefault('mistake_prob', 0.0)
    return g


def clip_names():
    """Goldens of ITAL(clip_cov = th) with more than 5 samples per batch (make_clip_golden.py)."""
    return [n for n in _npz_names() if n.startswith('clip_')]


def load_clip(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    g['updates'] = [dict(zip(g['upd%d_idx' % u].tolist(), g['upd%d_val' % u].tolist()))
                    for u in range(int(g['n_updates']))]
    g['steps'] = [dict(candidates=g['step%d_candidates' % t], mi=g['step%d_mi' % t], chosen=int(g['step%d_chosen' % t]))
                  for t in range(len(g['ret']))]
    return g


def baseline_names():
    """Goldens of EntropySampling / VarianceSampling (make_baseline_golden.py)."""
    return [n for n in _npz_names() if n.startswith('baseline_')]


def load_baseline(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    g['updates'] = [dict(zip(g['upd%d_idx' % u].tolist(), g['upd%d_val' % u].tolist()))
                    for u in range(int(g['n_updates']))]
    g['entropy_steps'] = []
    t = 0
    while 'entropy_step%d_entropy' % t in g:
        g['entropy_steps'].append(dict(candidates=g['entropy_step%d_candidates' % t], entropy=g['entropy_step%d_entropy' % t]))
        t += 1
    g['learner_kw'] = dict(length_scale=float(g['length_scale']), var=float(g['var']), noise=float(g['noise']))
    return g


def load_updpred(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    g['updates'] = [dict(zip(g['upd%d_idx' % u].tolist(), g['upd%d_val' % u].tolist()))
                    for u in range(int(g['n_updates']))]
    g['probes'] = [dict(feedback=dict(zip(g['probe%d_fb_idx' % p].tolist(), g['probe%d_fb_val' % p].tolist())),
                        test=g['probe%d_test' % p].tolist(), mean=g['probe%d_mean' % p], var=g['probe%d_var' % p],
                        cov=g['probe%d_cov' % p]) for p in range(int(g['n_probes']))]
    g['learner_kw'] = dict(length_scale=float(g['length_scale']), var=float(g['var']), noise=float(g['noise']))
    return g


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
