"""Drop-in proof: ital_b200.ITAL in the place of the reference's learner inside the reference's own experiment driver.

What `utils.LEARNERS['ITAL'] = ital_b200.ITAL` + `python run_experiment.py <config>` does is replayed
from a record of the UNMODIFIED reference run (/root/reference does not exist on the GPU box): the same constructor
keywords as utils.load_config passes (utils.py:110-119), then every reset / update / fetch_unlabelled / gp.predict
call of run_retrieval_experiment (run_experiment.py:133-168) in order.  Asserts: every batch identical to the
reference's, test-set predictions within 1e-6, and the `Round;Median_AP;...` table the driver prints reproduced
character for character (AP / NDCG to 1e-9)."""
import glob
import os

import numpy as np
import pytest

from dropin_replay import GOLDEN_DIR, load_experiment, replay, table_of

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(p)[len('experiment_'):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'experiment_*.npz')))


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('mode', ['default', 'streaming'])
def test_learner_replays_the_reference_experiment(name, mode):
    from ital_b200 import ITAL
    from oracle.ital_oracle import OracleITAL
    g = load_experiment(name)
    kw = {k: v for k, v in g['learner_kw'].items() if v is not None}
    learner = ITAL(g['X_train'], parallelized=True, lazy_rows=None if mode == 'default' else False, **kw)
    aps, ndcgs, ties = replay(learner, g, checker=OracleITAL(g['X_train'], **kw))
    assert table_of(aps, ndcgs) == g['table']
    assert ties == 0 or name.startswith('toy')          # only the toy set has structural ties
    # AP / NDCG against the oracle's replay of the same record, to 1e-9
    o_aps, o_ndcgs, _ = replay(OracleITAL(g['X_train'], **kw), g, checker=OracleITAL(g['X_train'], **kw))
    np.testing.assert_allclose(aps, o_aps, rtol=0, atol=1e-9)
    np.testing.assert_allclose(ndcgs, o_ndcgs, rtol=0, atol=1e-9)
    learner.close()
