"""Two-GPU run of the sharded learner over NCCL against the single-GPU learner and the oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _syn(n, d, seed, centres=20):
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((centres, d))
    assign = rng.integers(0, centres, n)
    X = C[assign] + 0.6 * rng.standard_normal((n, d))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    return X.astype(np.float32).astype(np.float64), assign


def _rounds(learner, y, rounds=3):
    learner.update({0: 1})
    out = []
    for _ in range(rounds):
        ret = learner.fetch_unlabelled(4)
        out.append(ret)
        learner.update({i: int(y[i]) for i in ret})
    return out, np.array(learner.rel_mean), (learner.top_results(), learner.top_results(25))


def _worker(rank, world, port, q, local_rows, peer):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['ITAL_B200_PEER'] = '1' if peer else '0'
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from ital_b200 import ITAL
        X, assign = _syn(5003, 96, seed=5)
        y = np.where(assign == assign[0], 1, -1)
        if local_rows:
            from ital_b200.dist import partition_rows
            off = partition_rows(len(X), world)
            learner = ITAL(X[off[rank]:off[rank + 1]], length_scale=1.0, device=rank, process_group=True,
                           local_rows=(int(off[rank]), len(X)))
        else:
            learner = ITAL(X, length_scale=1.0, device=rank, process_group=True)
        batches, rel_mean, tops = _rounds(learner, y)
        q.put((rank, batches, rel_mean, tops, learner._peer))
        learner.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('local_rows,peer', [(False, True), (True, True), (False, False)])
def test_two_gpus_match_one_gpu_and_oracle(local_rows, peer):
    """peer: proposals exchanged by peer stores over NVLink (ital_fetch_peer); otherwise the NCCL all-gather loop."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from ital_b200 import ITAL
    from oracle.ital_oracle import OracleITAL
    X, assign = _syn(5003, 96, seed=5)
    y = np.where(assign == assign[0], 1, -1)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, local_rows, peer)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    one_batches, one_mean, one_tops = _rounds(ITAL(X, length_scale=1.0, device=0), y)
    ora_batches, ora_mean, _ = _rounds(OracleITAL(X, length_scale=1.0), y)
    for rank, batches, rel_mean, tops, used_peer in res:
        assert used_peer == peer, 'peer exchange %s' % ('not available' if peer else 'not disabled')
        want = np.lexsort((np.arange(len(rel_mean)), -rel_mean))
        assert np.array_equal(tops[0], want) and np.array_equal(tops[1], want[:25])
        assert np.array_equal(one_tops[1], np.lexsort((np.arange(len(one_mean)), -one_mean))[:25])
        assert batches == one_batches == ora_batches
        np.testing.assert_allclose(rel_mean, one_mean, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(rel_mean, ora_mean, rtol=1e-6, atol=1e-9)


MODES = {'subset': (dict(length_scale=1.0, change_estimation_subset=3), 3),
         'subset_mistakes': (dict(length_scale=1.0, change_estimation_subset=2, mistake_prob=0.2), 3),
         'clip_cov': (dict(length_scale=1.0, clip_cov=0.3), 7)}


def _mode_worker(rank, world, port, q, mode):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from ital_b200 import ITAL
        X, fb = _mode_problem()
        kw, k = MODES[mode]
        learner = ITAL(X, device=rank, process_group=True, **kw)
        learner.update(fb)
        np.random.seed(21)                              # every rank draws the same subset
        ret = learner.fetch_unlabelled(k)
        q.put((rank, ret, getattr(learner, 'last_subset', None), list(learner.last_fetch_scores)))
        learner.close()
    finally:
        dist.destroy_process_group()


def _mode_problem(n=61, seed=2):
    rng = np.random.RandomState(seed)
    X = rng.randn(n, 2) * 1.3
    y = np.where(X[:, 0] - 0.4 * X[:, 1] > 0, 1, -1)
    return X, {1: int(y[1]), 6: int(y[6]), 20: int(y[20]), 40: int(y[40])}


@pytest.mark.parametrize('mode', sorted(MODES))
def test_subset_and_clip_modes_on_two_gpus(mode):
    """change_estimation_subset (stepwise protocol: records of batch + subset summed over the ranks, the shards' best
    records gathered) and clip_cov (per-shard adjacency pass and node sets) over two shards: same subset, batch and
    scores as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from ital_b200 import ITAL
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mode_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    X, fb = _mode_problem()
    kw, k = MODES[mode]
    one = ITAL(X, device=0, **kw)
    one.update(fb)
    np.random.seed(21)
    want = one.fetch_unlabelled(k)
    for rank, ret, subset, scores in res:
        assert subset == getattr(one, 'last_subset', None) and ret == want
        np.testing.assert_allclose(scores, one.last_fetch_scores, rtol=1e-9, atol=1e-12)
