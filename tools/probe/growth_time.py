"""Wall time of the fetch in which the projection matrix doubles its column capacity (one-off per doubling)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import bench
from ital_b200 import ITAL

X, assign = bench.syn_block(0, 1000000, 512)
L = ITAL(X, length_scale=1.0)
for fb in bench.labelled_state(assign[:65536]):
    L.update(fb)
nxt = 1000
out = []
for rnd in range(6):
    t0 = time.perf_counter()
    b = L.fetch_unlabelled(4)
    t1 = time.perf_counter()
    L.update({nxt + k: (1 if assign[nxt + k] == assign[0] else -1) for k in range(4)})
    nxt += 4
    import torch
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out.append('W=%d fetch %.2f ms update %.2f ms' % (int(L._shard.lib.ital_width(L._shard.handle)), (t1 - t0) * 1e3, (t2 - t1) * 1e3))
print(' | '.join(out))
