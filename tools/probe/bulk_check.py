import numpy as np, sys, time
sys.path.insert(0, '.')
from ital_b200 import ITAL
import bench
X, assign = bench.syn_block(0, 200000, 512)
fbs = bench.labelled_state(assign[:65536])
res = {}
for bulk in (False, True):
    L = ITAL(X, length_scale=1.0, bulk_stream=bulk)
    for fb in fbs: L.update(fb)
    ret = L.fetch_unlabelled(4)
    res[bulk] = (ret, np.array(L.last_fetch_scores), L.rel_mean.copy())
    print('bulk', bulk, ret, L.last_fetch_scores)
print('same batch', res[False][0] == res[True][0], 'scores bitwise', np.array_equal(res[False][1], res[True][1]), 'rel_mean bitwise', np.array_equal(res[False][2], res[True][2]))
