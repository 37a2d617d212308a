"""Small workloads for ncu captures on SYN-1M (tools/profile_r2.sh):  python tools/probe/prof_driver.py MODE
MODE: fused | exhaustive | general | update | streaming"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from ital_b200 import ITAL  # noqa: E402

mode = sys.argv[1]
n = int(os.environ.get('ROWS', 1000000))
X, assign = bench.syn_block(0, n, 512)
L = ITAL(X, length_scale=1.0)
for fb in bench.labelled_state(assign[:65536]):
    L.update(fb)
if mode == 'fused':
    for _ in range(4):
        L.fetch_unlabelled(4)
elif mode == 'exhaustive':
    L.exhaustive = True
    L.fetch_unlabelled(4)
elif mode == 'general':
    L.label_prob = 0.25
    L.fetch_unlabelled(4)
elif mode == 'streaming':
    L.lazy_rows = False
    for _ in range(2):
        L.fetch_unlabelled(4)
elif mode == 'update':
    c0 = assign[0]
    for _ in range(3):
        ret = L.fetch_unlabelled(4)
        L.update({i: (1 if assign[i] == c0 else -1) for i in ret})
print('done', mode)
