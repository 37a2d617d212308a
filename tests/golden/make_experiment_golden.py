"""Golden records of the reference's own experiment driver, for the drop-in test of the learner.

    python tests/golden/make_experiment_golden.py [case ...]

Runs the UNMODIFIED `run_experiment.run_retrieval_experiment` (/root/reference/run_experiment.py:79-212) with the
learner `utils.load_config` builds from a config file (/root/reference/utils.py:85-121, LEARNERS['ITAL']) through
oracle/ref_shims (numpy numexpr stub, mvndst stand-in, matplotlib / skimage stubs), and records every call the driver
makes on the learner -- reset(), update(feedback), fetch_unlabelled(k) and gp.predict(X_test) -- with its results,
plus the per-round AP / NDCG lists and the `Round;Median_AP;...` table the driver prints.  The GPU test
(tests/test_gpu_dropin.py) replays the same call sequence on ital_b200.ITAL: /root/reference does not exist on the
GPU box, only this file's output travels.

Outputs: tests/golden/experiment_<case>.npz
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

ital = ref_shims.install(with_plot_stubs=True)
REF = ref_shims.REFERENCE_ROOT


def record_experiment(name, config_file, overrides, learner_overrides=None):
    os.chdir(REF)                       # config files name their data relative to the reference root
    import run_experiment
    import utils
    from ital.gp import GaussianProcess
    config, dataset, learner = utils.load_config(config_file, 'EXPERIMENT', dict(overrides))
    for k, v in (learner_overrides or {}).items():
        setattr(learner, k, v)
    log = []
    arrays = {}

    def put(arr):
        key = 'a%d' % len(arrays)
        arrays[key] = np.asarray(arr)
        return key

    orig = dict(reset=learner.reset, update=learner.update, fetch=learner.fetch_unlabelled, predict=GaussianProcess.predict)

    def reset():
        log.append(dict(op='reset'))
        return orig['reset']()

    def update(feedback):
        log.append(dict(op='update', idx=[int(i) for i in feedback.keys()], val=[float(v) for v in feedback.values()]))
        return orig['update'](feedback)

    def fetch(k, *a, **kw):
        ret = orig['fetch'](k, *a, **kw)
        log.append(dict(op='fetch', k=int(k), ret=[int(i) for i in ret]))
        return ret

    def predict(self, X, *a, **kw):
        out = orig['predict'](self, X, *a, **kw)
        log.append(dict(op='predict', out=put(out)))
        return out

    learner.reset, learner.update, learner.fetch_unlabelled = reset, update, fetch
    GaussianProcess.predict = predict
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            run_experiment.run_retrieval_experiment(config, dataset, learner)
    finally:
        GaussianProcess.predict = orig['predict']
    table = [ln for ln in buf.getvalue().splitlines() if ln and (ln[0].isdigit() or ln.startswith('Round;'))]
    lbls = str(config.get('EXPERIMENT', 'query_classes', fallback='')).split() or list(dataset.class_relevance.keys())
    rel = {str(l): dataset.class_relevance[type(list(dataset.class_relevance.keys())[0])(l)] for l in lbls}
    kw = {k: getattr(learner, k) for k in ('length_scale', 'var', 'noise', 'label_prob', 'mistake_prob', 'top_candidates',
                                            'label_estimation')}
    out = dict(X_train=np.asarray(dataset.X_train_norm, dtype=np.float64), X_test=np.asarray(dataset.X_test_norm, dtype=np.float64),
               log=np.array(json.dumps(log)), table=np.array('\n'.join(table)), learner_kw=np.array(json.dumps(kw)),
               classes=np.array(json.dumps([str(l) for l in lbls])),
               config=np.array(json.dumps(dict(config_file=os.path.relpath(config_file, REF) if config_file.startswith(REF)
                                                else os.path.basename(config_file), overrides=overrides))))
    for l, (r_train, r_test) in rel.items():
        out['rel_train_' + l] = np.asarray(r_train)
        out['rel_test_' + l] = np.asarray(r_test)
    out.update(arrays)
    np.savez_compressed(os.path.join(HERE, 'experiment_%s.npz' % name), **out)
    print('%-24s %d calls recorded; table:\n%s' % (name, len(log), '\n'.join(table)), flush=True)


def butterflies_subset_conf(tmp, n_train=160, n_test=80, seed=3):
    """A small StoredDataset cut from data/butterflies_pca50.npz (the reference's n-by-n design and the Python MI loop
    make the full 1000-row experiment a matter of hours) and a config like configs/butterflies.conf for it."""
    d = np.load(os.path.join(REF, 'data', 'butterflies_pca50.npz'))
    rng = np.random.default_rng(seed)
    tr = np.sort(rng.choice(len(d['X_train']), n_train, replace=False))
    te = np.sort(rng.choice(len(d['X_test']), n_test, replace=False))
    data = os.path.join(tmp, 'butterflies_sub.npz')
    np.savez(data, X_train=d['X_train'][tr], y_train=d['y_train'][tr], X_test=d['X_test'][te], y_test=d['y_test'][te])
    conf = os.path.join(tmp, 'butterflies_sub.conf')
    with open(conf, 'w') as f:
        f.write('[EXPERIMENT]\ndataset = Stored\navg_class_perf = yes\nmethod = ITAL\nbatch_size = 4\nrounds = 3\n'
                'repetitions = 2\nlabel_prob = 1.0\nmistake_prob = 0.0\nquery_classes = 0 3\n\n[Stored]\ndata_file = %s\n\n'
                '[METHOD_DEFAULTS]\nlength_scale = 2.5\n\n[ITAL]\nlabel_prob = 1.0\nmistake_prob = 0.0\n' % data)
    return conf


if __name__ == '__main__':
    want = set(sys.argv[1:])
    tmp = tempfile.mkdtemp()
    cases = [
        # configs/toy-demo.conf with the batch size BASELINE.json quotes for the toy config (the file ships 6)
        ('toy_demo_b4', os.path.join(REF, 'configs', 'toy-demo.conf'), {'batch_size': '4'}),
        ('butterflies_sub', butterflies_subset_conf(tmp), {}),
        # as shipped: batch_size = 6 (five and six variables go through the stand-in at its lower orders)
        ('toy_demo', os.path.join(REF, 'configs', 'toy-demo.conf'), {}),
    ]
    for name, conf, ov in cases:
        if want and name not in want:
            continue
        record_experiment(name, conf, ov)
