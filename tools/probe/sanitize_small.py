"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once."""
import sys
import numpy as np
sys.path.insert(0, '.')
from ital_b200 import ITAL

rng = np.random.default_rng(0)
for d, storage in ((512, 'float32'), (70, 'float64')):
    X = rng.standard_normal((3000, d))
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X = X.astype(np.float32).astype(np.float64)
    for kw in (dict(), dict(lazy_rows=True), dict(bulk_stream=False), dict(label_prob=0.6, mistake_prob=0.1),
               dict(mistake_prob=0.3), dict(exhaustive=True)):
        L = ITAL(X, length_scale=1.0, storage=storage, **kw)
        L.update({0: 1})
        L.update({5: -1, 9: 1, 11: -1, 40: -1, 41: 1})
        k = 3 if kw.get('label_prob', 1) < 1 else 6
        ret = L.fetch_unlabelled(k)
        L.update({i: (1 if j % 2 else -1) for j, i in enumerate(ret)})
        L.gp.predict(X[:50], cov_mode='diag')
        print(d, storage, kw, ret, flush=True)
print('sanitize run ok')
