"""Per-step timeline of the fused persistent fetch kernel on N GPUs (torchrun): where a greedy step's time goes on every
rank -- the local phases, and the commit phase, which in the multi-GPU kernel contains the peer exchange (stores into
the peers' buffers over NVLink, wait for every peer's flag) and with it the skew between the ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/probe/fused_trace_multi.py
"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ital_b200 import ITAL  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rows = int(os.environ.get('ROWS', 1000000))
X, assign = bench.syn_block(rank * rows, rows, 512)
head = assign[:65536] if rank == 0 else bench.syn_block(0, 65536, 512)[1]
L = ITAL(X, length_scale=1.0, device=local, process_group=True, local_rows=(rank * rows, rows * world))
for fb in bench.labelled_state(head):
    L.update(fb)
lib, h = L._shard.lib, L._shard.handle
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
names = ['S0 scan', 'S0 barrier', 'commit0+exchange']
for t in (1, 2, 3):
    names += ['P1 t%d' % t, 'bar', 'P2', 'bar', 'P3 masses', 'P3 eval', 'bar', 'P4', 'bar', 'P5', 'bar', 'P6', 'commit+exchange']
acc = []
for rep in range(14):
    flush.fill_(1)
    torch.cuda.synchronize()
    dist.barrier()
    lib.ital_fused_trace(h, 1, None, 0)
    ret = L.fetch_unlabelled(4)
    out = (ctypes.c_uint64 * 64)()
    lib.ital_fused_trace(h, 1, out, 64)
    st = np.array(out[:len(names) + 1], dtype=np.float64)
    if rep >= 2:
        acc.append(np.diff(st) / 1e3)
d = np.median(np.array(acc), axis=0)
t = torch.tensor(d, device='cuda')
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
if rank == 0:
    A = np.array([x.cpu().numpy() for x in allt])            # (world, phases)
    print('# fused fetch_unlabelled(4), %d GPUs x %d rows, L2 flushed, peer exchange %s; medians of 12 fetches, us, CTA 0 of every rank'
          % (world, rows, bool(L._peer)))
    print('%-18s %8s %8s %8s' % ('phase', 'min', 'median', 'max'))
    for k, nm in enumerate(names):
        print('%-18s %8.1f %8.1f %8.1f' % (nm, A[:, k].min(), np.median(A[:, k]), A[:, k].max()))
    tot = A.sum(axis=1)
    ex = A[:, [k for k, nm in enumerate(names) if 'exchange' in nm]].sum(axis=1)
    print('total per rank: min %.1f median %.1f max %.1f us; commit+exchange phases per rank: min %.1f median %.1f max %.1f us'
          % (tot.min(), np.median(tot), tot.max(), ex.min(), np.median(ex), ex.max()))
    print(json.dumps({'world': world, 'batch': [int(i) for i in ret], 'total_us_median': float(np.median(tot)),
                      'exchange_us_median': float(np.median(ex))}))
L.close()
dist.destroy_process_group()
