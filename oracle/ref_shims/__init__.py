"""Shims that let the UNMODIFIED reference (/root/reference, read-only) run in this container.

Test infrastructure only (golden-vector generation); nothing here ships or runs on the GPU box.

* ``numexpr`` -> numpy (see numexpr/__init__.py).
* ``scipy.stats.mvn.mvndst`` was removed from scipy; ``install()`` injects a deterministic, high-order
  stand-in with the same signature and return convention ``(error, value, inform)``.
* ``matplotlib`` / ``skimage`` stubs so that ``run_experiment.py`` can be imported (viz_utils.py:3-4).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'
# Gauss-Legendre nodes per panel of the stand-in, by number of variables (the oracle itself uses 32/16/12
# for 2/3/4 variables); two variables go to scipy's own translation of Genz's BVU instead.
STANDIN_Q = {3: 48, 4: 28, 5: 14, 6: 12}
STANDIN_R = 8.5


def mvndst_standin(lower, upper, infin, correl, maxpts=None, abseps=None, releps=None):
    """Same contract as Genz's MVNDST for the calls the reference makes (ital.py:380, 405, 425).

    Standardised limits, INFIN[i] = 1 -> [lower_i, inf), 0 -> (-inf, upper_i]; CORREL is the strict lower
    triangle in np.tril_indices order.  Evaluated with the oracle's nested Gauss-Legendre rule at a much
    higher order than the oracle uses, so that goldens are converged to ~1e-12 for well-conditioned blocks.
    """
    from oracle.orthant import orthant_prob_all
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    infin = np.asarray(infin)
    D = len(lower)
    corr = np.eye(D)
    i, j = np.tril_indices(D, -1)
    corr[i, j] = correl
    corr[j, i] = correl
    if np.any((infin != 0) & (infin != 1)):
        raise NotImplementedError('stand-in covers half-infinite limits only')
    pivot = np.where(infin == 1, lower, upper)
    if not np.all(np.isfinite(pivot)) or not np.all(np.isfinite(corr)):
        return 0.0, float('nan'), 0
    if np.all(np.abs(pivot) > 12.0):
        # every variable is more than 12 standard deviations from its limit: the probability is 0 or 1 to 1e-30 whatever
        # the correlations (the updated distributions of a user who labels everything, ital.py:432-450)
        inside = np.where(infin == 1, pivot < 0, pivot > 0)
        return 1e-30, float(np.all(inside)), 0
    if D == 2:
        from scipy.stats._qmvnt import _bvnu        # Genz BVU: P(x > h, y > k) for correlation r
        sgn = np.where(infin == 1, 1.0, -1.0)       # flip the variables bounded from above
        return 1e-15, float(_bvnu(sgn[0] * pivot[0], sgn[1] * pivot[1], sgn[0] * sgn[1] * corr[0, 1])), 0
    p = orthant_prob_all(-pivot, corr, q=STANDIN_Q.get(D, 6), R=STANDIN_R)
    idx = int(sum(int(b) << k for k, b in enumerate(infin)))
    return 1e-12, float(p[idx]), 0


def install(with_plot_stubs=False):
    """Put the shims and the reference on sys.path / into scipy; returns the imported ``ital`` package."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)                      # makes ``import numexpr`` resolve to the stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    import scipy.stats
    if not hasattr(scipy.stats, 'mvn') or not hasattr(scipy.stats.mvn, 'mvndst'):
        if not hasattr(scipy.stats, 'mvn'):
            scipy.stats.mvn = types.ModuleType('scipy.stats.mvn')
        scipy.stats.mvn.mvndst = mvndst_standin
    if with_plot_stubs:
        for name in ('matplotlib', 'matplotlib.pyplot', 'skimage', 'skimage.transform', 'skimage.io'):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
    import ital
    return ital
