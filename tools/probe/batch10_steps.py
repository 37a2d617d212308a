"""Per-step wall time of fetch_unlabelled(10) on SYN-1M through the stepwise API (host-synchronous per step)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import torch  # noqa: E402
from ital_b200 import ITAL  # noqa: E402

n = int(os.environ.get('ROWS', 1000000))
X, assign = bench.syn_block(0, n, 512)
L = ITAL(X, length_scale=1.0)
for fb in bench.labelled_state(assign[:65536]):
    L.update(fb)
sh = L._shard
for rep in range(3):
    L._apply_lazy_rows()
    sh.fetch_begin(1.0, 0.0)
    times, stats = [], []
    for it in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rec = sh.fetch_propose(-np.inf, False)
        t1 = time.perf_counter()
        sh.fetch_commit(rec)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        times.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
        stats.append(sh.stats()[:3].tolist())
    sh.fetch_end()
    print('rep', rep)
    for it, (tm, st) in enumerate(zip(times, stats)):
        print('  step %d propose %.2f ms commit %.2f ms  worklist/scored/nodes %s' % (it, tm[0], tm[1], st))
