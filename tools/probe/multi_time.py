"""Time the multi-column labelled pass (k_extend_bulk_multi / k_extend_multi) alone: 1M x 512 float32 pool,
repeated update() calls of 4 labels each with the library's per-launch CUDA events on."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, '.')
import bench
from ital_b200 import ITAL

X, assign = bench.syn_block(0, 1000000, 512)
L = ITAL(X, length_scale=1.0)
for fb in bench.labelled_state(assign[:65536]):
    L.update(fb)
lib, h = L._shard.lib, L._shard.handle
out = []
nxt = 1000
for rnd in range(12):
    lab = {nxt + k: (1 if assign[nxt + k] == assign[0] else -1) for k in range(4)}
    nxt += 4
    lib.ital_profile_enable(h, 1)
    L.update(lab)
    ms, nl, nb = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    lib.ital_profile_read(h, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(nb))
    lib.ital_profile_enable(h, 0)
    out.append((int(lib.ital_width(h)), ms.value / max(1, nl.value), nb.value / 1e9 / (ms.value / 1e3)))
print(' '.join('W=%d:%.3fms/%.0fGB/s' % o for o in out))
print('median ms %.4f' % np.median([o[1] for o in out]), 'rel_mean checksum %.12f' % float(np.sum(L.rel_mean)))
