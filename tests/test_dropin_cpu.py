"""Drop-in proof on the CPU side: the oracle, driven through the call sequence the reference's experiment driver made
on the reference's own learner (recorded by tests/golden/make_experiment_golden.py from the unmodified
run_experiment.run_retrieval_experiment + utils.load_config), selects the same batches and reproduces the printed
`Round;Median_AP;...` table.  tests/test_gpu_dropin.py does the same with ital_b200.ITAL."""
import glob
import os

import pytest

from dropin_replay import GOLDEN_DIR, load_experiment, replay, table_of
from oracle.ital_oracle import OracleITAL

CASES = sorted(os.path.basename(p)[len('experiment_'):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'experiment_*.npz')))


@pytest.mark.parametrize('name', CASES)
def test_oracle_replays_the_reference_experiment(name):
    g = load_experiment(name)
    kw = {k: v for k, v in g['learner_kw'].items() if v is not None}
    learner = OracleITAL(g['X_train'], **kw)
    aps, ndcgs, ties = replay(learner, g, checker=OracleITAL(g['X_train'], **kw))
    assert table_of(aps, ndcgs) == g['table']
    assert ties == 0 or name.startswith('toy')          # only the toy set has structural ties
