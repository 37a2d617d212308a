"""Host-side logic of the drop-in learner (ital_b200/learner.py) without a GPU: the shard (everything behind the C
ABI) is replaced by a recording stub, so that what is checked is the reference's learner state machine
(ital/retrieval_base.py:34-194, ital/ital.py:84-134): feedback partitioning and its errors, the order in which
labelled points enter the model, clamping of k, the top_candidates restriction, unnameable samples, reset."""
import types

import numpy as np
import pytest

from ital_b200 import learner as learner_mod


class StubLib(object):
    def __getattr__(self, name):
        return lambda *a: 0


class StubShard(object):
    """Stands in for learner._Shard: keeps the calls and answers with simple deterministic values."""

    def __init__(self, X, dtype_code, row_offset, n_data, length_scale, var, noise, device):
        self.n_local, self.d = X.shape
        self.n_data = n_data
        self.lib, self.handle = StubLib(), 1
        self.calls = []
        self.labelled, self.seen, self.restricted = [], set(), None
        self.mean = np.linspace(-1.0, 1.0, self.n_local)

    def close(self):
        self.calls.append(('close',))

    def reset(self):
        self.calls.append(('reset',))
        self.labelled, self.seen, self.restricted = [], set(), None

    def record_doubles(self):
        return 8 + 64 + self.d

    def export_points(self, idx):
        rec = np.zeros((len(idx), self.record_doubles()))
        rec[:, 0] = idx
        return rec

    def add_labelled_many(self, records, y):
        idx = [int(r[0]) for r in np.atleast_2d(records)]
        self.calls.append(('add', idx, [float(v) for v in y]))
        self.labelled += idx
        self.seen.update(idx)

    def update_labelled(self, idx, y):
        idx = [int(i) for i in idx]
        self.calls.append(('add', idx, [float(v) for v in y]))
        self.labelled += idx
        self.seen.update(idx)

    def mark_seen(self, idx):
        self.calls.append(('seen', [int(i) for i in idx]))
        self.seen.update(int(i) for i in idx)

    def restrict_candidates(self, idx):
        self.restricted = None if idx is None else set(int(i) for i in idx)
        self.calls.append(('restrict', None if idx is None else sorted(self.restricted)))

    def restrict_top(self, top):
        cand = [i for i in range(self.n_data) if i not in self.seen]
        order = sorted(cand, key=lambda i: (-self.mean[i], i))[:int(top)]
        self.restricted = set(order)
        self.calls.append(('restrict', sorted(self.restricted)))

    # -- stepwise protocol (change_estimation_subset) --
    def set_sub_mode(self, on):
        self.calls.append(('sub_mode', bool(on)))

    def fetch_begin(self, label_prob, mistake_prob):
        self.committed = []

    def fetch_commit(self, record):
        self.committed.append(int(record[0]))

    def fetch_end(self):
        pass

    def fetch_propose_sub(self, n_batch, only_row):
        self.calls.append(('propose_sub', n_batch, list(self.committed), int(only_row)))
        rec = np.zeros(self.record_doubles())
        cand = [i for i in range(self.n_data) if i not in self.seen and i not in self.committed
                and (self.restricted is None or i in self.restricted)]
        if only_row >= 0:
            cand = [i for i in cand if i == only_row]
        if not cand:
            rec[0] = -1
            return rec
        score = {i: -abs(i - 10.2) for i in cand}           # peaks at row 10, then 11, 9, 12, ...
        best = max(cand, key=lambda i: (score[i], -i))
        rec[0], rec[1] = best, score[best]
        return rec

    def last_scores(self):
        return np.full(self.n_local, np.nan)

    def fetch(self, k, label_prob, mistake_prob, exhaustive):
        cand = [i for i in range(self.n_data) if i not in self.seen
                and (self.restricted is None or i in self.restricted)]
        self.calls.append(('fetch', k, label_prob, mistake_prob, bool(exhaustive)))
        return np.array(cand[:k], dtype=np.int64), np.arange(len(cand[:k]), dtype=np.float64)

    def stats(self):
        return np.zeros(8)

    def rel_mean(self):
        return self.mean.copy()

    def rel_var(self):
        return np.full(self.n_local, 0.5)

    def top_results(self, k):
        order = np.lexsort((np.arange(self.n_data), -self.mean[:self.n_data]))
        return (order if k is None else order[:k]), self.mean[order if k is None else order[:k]]


@pytest.fixture
def make(monkeypatch):
    monkeypatch.setattr(learner_mod, '_Shard', StubShard)

    def _make(n=20, d=3, **kw):
        X = np.arange(n * d, dtype=np.float64).reshape(n, d)
        return learner_mod.ITAL(X, length_scale=1.0, **kw)
    return _make


def test_update_orders_relevant_before_irrelevant_in_chunks_of_four(make):
    L = make()
    L.update({5: -1, 3: 1, 9: -2.0, 7: 0.5, 11: 1, 2: -1, 13: 0})
    adds = [c for c in L._shard.calls if c[0] == 'add']
    # retrieval_base.py:116-119: rel + irr, each in dict order; here at most four points per pass
    assert [c[1] for c in adds] == [[3, 7, 11, 5], [9, 2]]
    assert [c[2] for c in adds] == [[1.0, 1.0, 1.0, -1.0], [-1.0, -1.0]]
    assert L.relevant_ids == {3, 7, 11} and L.irrelevant_ids == {5, 9, 2} and L.unnameable_ids == {13}
    assert ('seen', [13]) in L._shard.calls and L.rounds == 1
    assert L.gp.ind == [3, 7, 11, 5, 9, 2] and L.gp.y.tolist() == [1, 1, 1, -1, -1, -1]


def test_feedback_cannot_change_and_repeats_are_ignored(make):
    L = make()
    L.update({4: 1, 6: -1})
    with pytest.raises(RuntimeError, match='Cannot change feedback once given.'):
        L.update({4: -1})
    with pytest.raises(RuntimeError, match='Cannot change feedback once given.'):
        L.update({6: 1})
    n_calls = len(L._shard.calls)
    L.update({4: 1, 6: -3})                            # same labels again: nothing happens, no round counted
    assert len(L._shard.calls) == n_calls and L.rounds == 1
    L.update({8: 0})                                   # unnameable only: seen, but no round (retrieval_base.py:121-126)
    assert L.rounds == 1 and L.unnameable_ids == {8}


def test_fetch_needs_a_label_clamps_k_and_passes_the_feedback_model(make):
    L = make(n=6, label_prob=0.25, mistake_prob=0.1, exhaustive=True)
    assert L.rel_mean is None
    with pytest.raises(RuntimeError):
        L.fetch_unlabelled(2)
    L.update({0: 1, 1: -1, 2: 0})
    assert L.fetch_unlabelled(10) == [3, 4, 5]         # k clamped to the unseen rows (ital.py:99-100)
    assert ('fetch', 3, 0.25, 0.1, True) in L._shard.calls
    assert L.get_unseen() == [3, 4, 5]
    assert L.fetch_unlabelled(0) == []


def test_top_candidates_restricts_to_the_best_means_and_lifts_the_restriction(make):
    L = make(n=12, top_candidates=4)
    L.update({11: 1})                                  # the stub's means grow with the row index
    ret = L.fetch_unlabelled(2)
    restricts = [c for c in L._shard.calls if c[0] == 'restrict']
    assert restricts == [('restrict', [7, 8, 9, 10]), ('restrict', None)]      # ital.py:111-117, then undone
    assert ret == [7, 8]
    F = make(n=12, top_candidates=2.0)                 # float: multiple of the labelled + query count
    F.update({11: 1, 10: -1})
    F.fetch_unlabelled(1)
    assert [c for c in F._shard.calls if c[0] == 'restrict'][0] == ('restrict', [6, 7, 8, 9])


def test_queries_are_fitted_as_relevant_rows_behind_the_pool(make):
    L = make(n=5, queries=[np.zeros(3), np.ones(3)])
    assert L._shard.n_local == 7 and L._shard.n_data == 5
    assert [c for c in L._shard.calls if c[0] == 'add'] == [('add', [5, 6], [1.0, 1.0])]    # retrieval_base.py:40,57
    assert len(L.rel_mean) == 5 and L.fetch_unlabelled(2) == [0, 1]
    assert L.top_results(3).tolist() == [4, 3, 2] and L.top_results().tolist() == [4, 3, 2, 1, 0]
    assert L.top_results(-2).tolist() == [4, 3, 2] and len(L.top_results(0)) == 0           # ind[:k] semantics


def test_reset_and_unsupported_modes(make):
    L = make()
    L.update({1: 1})
    L.reset()
    assert L.rounds == 0 and L.rel_mean is None and L.get_unseen() == list(range(20))
    for kw in (dict(change_estimation_subset=None), dict(change_estimation_subset=3, label_prob=0.5),
               dict(change_estimation_subset=12)):
        B = make(**kw)
        B.update({0: 1})
        with pytest.raises(NotImplementedError):
            B.fetch_unlabelled(1)
    C = make(clip_cov=0.5)                             # ital.py:360: grouping only for more than 5 variables
    C.update({0: 1})
    assert len(C.fetch_unlabelled(5)) == 5
    assert len(C.fetch_unlabelled(6)) == 6
    G = make(clip_cov=0.5, label_prob=0.5)             # ... the general model stops at 5 samples with or without it
    G.update({0: 1})
    assert len(G.fetch_unlabelled(5)) == 5
    with pytest.raises(NotImplementedError):
        G.fetch_unlabelled(6)
    M = make(monte_carlo_num_rel=3, monte_carlo_num_fb=5)      # evaluated exactly (zero-variance limit of the estimator)
    M.update({0: 1})
    assert len(M.fetch_unlabelled(4)) == 4
    with pytest.raises(ValueError):
        make(storage='float16')


def test_bad_sample_indices_raise_index_error(make):
    """The reference indexes numpy arrays with the feedback keys (IndexError when out of range); here a row that no
    shard owns would otherwise enter the model as an all-zero record."""
    L = make(n=10)
    for bad in (10, -1, 250):
        with pytest.raises(IndexError):
            L.update({bad: 1})
    with pytest.raises(IndexError):
        L.update({3: 1, 12: 0})
    assert L.rounds == 0 and not [c for c in L._shard.calls if c[0] == 'add']     # nothing reached the model
    L.update({3: 1})
    with pytest.raises(IndexError):
        L.updated_prediction({11: 1}, [0, 1])


def test_batch_size_is_validated_before_any_work(make):
    """Batches beyond what the node sets cover are refused up front, as NotImplementedError (configs/toy-mistakes.conf
    ships batch_size = 6 with label_prob = 0.75), not in the middle of the greedy loop."""
    L = make(n=40, label_prob=0.75, mistake_prob=0.2)
    L.update({0: 1})
    n_calls = len(L._shard.calls)
    with pytest.raises(NotImplementedError, match='label_prob < 1'):
        L.fetch_unlabelled(6)
    assert len(L._shard.calls) == n_calls
    assert len(L.fetch_unlabelled(5)) == 5
    P = make(n=40)
    P.update({0: 1})
    with pytest.raises(NotImplementedError):
        P.fetch_unlabelled(12)
    assert len(P.fetch_unlabelled(11)) == 11


def test_change_estimation_subset_draws_like_the_reference_and_scores_members_separately(make):
    """ital.py:105-106: the subset is np.random.choice(candidates, c, replace=False), sorted, from the global RNG;
    per greedy step all candidates outside batch + subset are scored in one evaluation, every subset member that is
    still a candidate in one of its own with itself moved out of the subset (ital.py:516-519)."""
    L = make(n=20, change_estimation_subset=3)
    L.update({0: 1, 1: -1})
    np.random.seed(4)
    want_subset = sorted(int(i) for i in np.random.choice(list(range(2, 20)), 3, replace=False))
    np.random.seed(4)
    ret = L.fetch_unlabelled(2)
    assert L.last_subset == want_subset
    assert ret == [10, 11]                                  # the stub's scores peak at row 10, then 11
    props = [c for c in L._shard.calls if c[0] == 'propose_sub']
    S = want_subset
    first = [('propose_sub', 0, S, -1)] + [('propose_sub', 0, [j for j in S if j != i], i) for i in S]
    assert props[:4] == first
    sub2 = [i for i in S if i != 10]
    second = [('propose_sub', 1, [10] + sub2, -1)] + [('propose_sub', 1, [10] + [j for j in sub2 if j != i], i)
                                                      for i in sub2]
    assert props[4:] == second
    assert [c for c in L._shard.calls if c[0] == 'sub_mode'] == [('sub_mode', True), ('sub_mode', False)]
