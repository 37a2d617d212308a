"""Phase timeline of the fused persistent fetch kernel on SYN-1M (CTA 0's view, %globaltimer)."""
import ctypes
import os
import sys

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
lib, h = L._shard.lib, L._shard.handle
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
names = ['S0 scan', 'S0 barrier', 'commit0']
for t in (1, 2, 3):
    names += ['P1 t%d' % t, 'bar', 'P2', 'bar', 'P3 masses', 'P3 eval', 'bar', 'P4', 'bar', 'P5', 'bar', 'P6', 'commit']
for cold in (True, False):
    acc = []
    for rep in range(12):
        if cold:
            flush.fill_(1)
        torch.cuda.synchronize()
        lib.ital_fused_trace(h, 1, None, 0)
        L.fetch_unlabelled(4)
        out = (ctypes.c_uint64 * 64)()
        m = lib.ital_fused_trace(h, 1, out, 64)
        st = np.array(out[:len(names) + 1], dtype=np.float64)
        if rep >= 2:
            acc.append(np.diff(st) / 1e3)
    d = np.median(np.array(acc), axis=0)
    print('cold L2' if cold else 'warm L2', 'total %.1f us' % d.sum())
    for nm, v in zip(names, d):
        print('  %-12s %7.1f' % (nm, v))
