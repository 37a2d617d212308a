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
