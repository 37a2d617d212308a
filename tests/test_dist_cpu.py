"""Host-side multi-GPU plumbing on CPU: row partition, winner selection, and the record exchange over a
world_size-2 gloo group (the same TorchComm code path moves CUDA tensors over NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest

from ital_b200.dist import LocalComm, merge_top, partition_rows, pick_winner


def test_partition_rows_contiguous_and_balanced():
    for n, w in ((10, 3), (1000003, 8), (5, 5), (7, 1)):
        off = partition_rows(n, w)
        assert off[0] == 0 and off[-1] == n and len(off) == w + 1
        sizes = np.diff(off)
        assert sizes.max() - sizes.min() <= 1 and sizes.min() >= 0


def test_pick_winner_score_then_lowest_index():
    rec = np.zeros((4, 12))
    rec[:, 0] = [40, 7, 19, -1]
    rec[:, 1] = [0.5, 0.7, 0.7, -np.inf]
    assert pick_winner(rec) == 1                 # tie on score -> lower global row
    rec[1, 1] = np.nan
    assert pick_winner(rec) == 2                 # NaN never wins
    rec[:, 0] = -1
    assert pick_winner(rec) == -1                # all shards empty
    assert LocalComm().gather_records(rec[0]).shape == (1, 12)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from ital_b200.dist import TorchComm
        comm = TorchComm()
        n = 11
        off = partition_rows(n, world)
        lo, hi = off[rank], off[rank + 1]
        # export: only the owner of a row contributes a non-zero record
        rec = np.zeros((2, 9))
        for a, g in enumerate((3, 9)):
            if lo <= g < hi:
                rec[a] = np.arange(9) + 100 * g
        summed = comm.sum_records(rec)
        # propose: every shard offers its best candidate, everyone picks the same winner
        mine = np.zeros(9)
        mine[0] = lo + 1
        mine[1] = 0.25 if rank == 0 else 0.75
        allrec = comm.gather_records(mine)
        rows = comm.gather_rows(np.arange(lo, hi, dtype=np.float64), off)
        # top_results: every shard sends its own descending list, everyone merges to the same global list
        means = np.array([0.5, -1.0, 0.5, 2.0, 0.0, 2.0, -3.0, 0.25, 0.5, 1.0, -0.5])
        loc = np.arange(lo, hi)
        order = np.lexsort((loc, -means[lo:hi]))
        top_all = merge_top(comm, loc[order], means[lo:hi][order], int(np.max(np.diff(off))), None)
        top4 = merge_top(comm, loc[order][:4], means[lo:hi][order][:4], 4, 4)
        hb = comm.gather_bytes(np.full(16, rank + 1, dtype=np.uint8))
        assert hb.shape == (world, 16) and all(np.all(hb[r] == r + 1) for r in range(world))
        assert comm.all_agree(True) and not comm.all_agree(rank == 0)
        comm.barrier()
        out.put((rank, summed.tolist(), pick_winner(allrec), rows.tolist(), top_all.tolist(), top4.tolist()))
    finally:
        dist.destroy_process_group()


def test_record_exchange_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    means = np.array([0.5, -1.0, 0.5, 2.0, 0.0, 2.0, -3.0, 0.25, 0.5, 1.0, -0.5])
    want = np.lexsort((np.arange(11), -means)).tolist()
    for rank, summed, win, rows, top_all, top4 in res:
        assert top_all == want and top4 == want[:4]
        assert summed[0] == (np.arange(9) + 300).tolist() and summed[1] == (np.arange(9) + 900).tolist()
        assert win == 1
        assert rows == list(range(11))


class _StubLib(object):
    def __getattr__(self, name):
        return lambda *a: 0


class _StubShard(object):
    """Row-sharded stand-in for learner._Shard (no GPU): scores are a fixed function of the global row, so the batch
    the sharded learner must return is known in closed form."""

    def __init__(self, X, dtype_code, row_offset, n_data, length_scale, var, noise, device):
        self.n_local, self.d = X.shape
        self.lo, self.n_data = int(row_offset), int(n_data)
        self.lib, self.handle = _StubLib(), 1
        self.seen, self.selected, self.labelled = set(), set(), []

    @staticmethod
    def score(i):
        return ((i * 7919) % 101) / 101.0

    def close(self):
        pass

    def reset(self):
        self.seen, self.selected, self.labelled = set(), set(), []

    def record_doubles(self):
        return 8 + 64 + self.d

    def export_points(self, idx):
        rec = np.zeros((len(idx), self.record_doubles()))
        for a, g in enumerate(idx):
            if self.lo <= g < self.lo + self.n_local:
                rec[a, 0] = g
                rec[a, 2] = 100.0 + g                 # non-owners contribute zeros; the sum is the owner's record
        return rec

    def add_labelled_many(self, records, y):
        records = np.atleast_2d(records)
        assert all(r[2] == 100.0 + r[0] for r in records), 'incomplete record after the exchange'
        self.labelled += [int(r[0]) for r in records]
        self.seen.update(int(r[0]) for r in records)

    def mark_seen(self, idx):
        self.seen.update(int(i) for i in idx)

    def fetch_begin(self, label_prob, mistake_prob):
        self.selected = set()

    def fetch_propose(self, floor_score, exhaustive):
        rec = np.zeros(self.record_doubles())
        rec[0], rec[1] = -1.0, -np.inf
        for g in range(self.lo, min(self.lo + self.n_local, self.n_data)):
            if g in self.seen or g in self.selected:
                continue
            if self.score(g) > rec[1]:
                rec[0], rec[1] = g, self.score(g)
        return rec

    def fetch_commit(self, record):
        assert record[0] >= 0 and (record[2] == 100.0 + record[0] or record[2] == 0.0), 'incomplete record'
        self.selected.add(int(record[0]))

    # -- change_estimation_subset: batch + subset are committed as batch columns, then one proposal per evaluation --
    def set_sub_mode(self, on):
        self.sub_mode = bool(on)

    def fetch_propose_sub(self, n_batch, only_row):
        rec = np.zeros(self.record_doubles())
        rec[0], rec[1] = -1.0, -np.inf
        for g in range(self.lo, min(self.lo + self.n_local, self.n_data)):
            if g in self.seen or g in self.selected or (only_row >= 0 and g != only_row):
                continue
            sc = self.score(g) - 0.001 * n_batch
            if sc > rec[1]:
                rec[0], rec[1] = g, sc
        return rec

    def last_scores(self):
        return np.full(self.n_local, np.nan)

    def fetch_end(self):
        self.selected = set()

    def stats(self):
        return np.zeros(8)

    def rel_mean(self):
        return np.arange(self.lo, self.lo + self.n_local, dtype=np.float64)


def _learner_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from ital_b200 import learner as learner_mod
        learner_mod._Shard = _StubShard
        n = 23
        X = np.arange(n * 2, dtype=np.float64).reshape(n, 2)
        L = learner_mod.ITAL(X, length_scale=1.0, process_group=True)
        L.update({3: 1, 20: -1, 11: 0})                # rows 3 and 20 live on different shards
        batch = L.fetch_unlabelled(5)
        # change_estimation_subset over the same two shards: every rank draws the same subset (same seed), the records
        # of batch + subset are completed over the ranks before they are committed, proposals are gathered per evaluation
        S = learner_mod.ITAL(X, length_scale=1.0, process_group=True, change_estimation_subset=3)
        S.update({3: 1, 20: -1, 11: 0})
        np.random.seed(5)
        sub_batch = S.fetch_unlabelled(3)
        out.put((rank, batch, L._shard.labelled, L.rel_mean.tolist(), int(L._shard.n_local), bool(L._peer),
                 sub_batch, list(S.last_subset)))
        S.close()
        L.close()
    finally:
        dist.destroy_process_group()


def test_sharded_learner_host_loop_over_gloo_world2():
    """The learner's multi-shard host logic end to end on CPU: contiguous row blocks, summed point records for the
    labelled rows, per-step gather of the shards' proposals and the same winner everywhere (ital.py:119-132)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_learner_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cand = [i for i in range(23) if i not in (3, 20, 11)]
    want = sorted(cand, key=lambda i: (-_StubShard.score(i), i))[:5]
    assert sorted(r[4] for r in res) == [11, 12]
    np.random.seed(5)
    want_subset = sorted(int(i) for i in np.random.choice(cand, 3, replace=False))
    for rank, batch, labelled, rel_mean, n_local, peer, sub_batch, subset in res:
        assert subset == want_subset
        assert sub_batch == want[:3]                   # (the stub's scores do not depend on the subset)
        assert batch == want
        assert labelled == [3, 20]
        assert rel_mean == list(map(float, range(23)))
        assert not peer                                # no CUDA, no peer exchange: the host loop is what ran
