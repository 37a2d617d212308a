import numpy as np, sys
sys.path.insert(0, '.')
from ital_b200 import ITAL
for n in (1000000, 1200000):
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, 8)).astype(np.float32)
    X[n // 2] = X[3]
    L = ITAL(X, length_scale=2.0)
    L.update({1: -1, 2: 1, 7: 1})
    rm = np.array(L.rel_mean)
    want = np.lexsort((np.arange(n), -rm))
    got = L.top_results()
    print(n, 'full', np.array_equal(got, want), 'top100', np.array_equal(L.top_results(100), want[:100]))
    L.close()
