"""Deterministic Gaussian orthant probabilities -- ORACLE (test infrastructure, not product code).

The reference evaluates ``P(sign(z) = r)``, ``z ~ N(mean, cov)``, with ``scipy.stats.norm.cdf`` for one
variable (/root/reference/ital/ital.py:364-369) and with Genz's Fortran ``MVNDST`` for two or more
(/root/reference/ital/ital.py:373-383, ``maxpts = 100*dim, abseps = releps = 1e-4``).  ``MVNDST`` is a
third-party routine that used to ship inside scipy (README pins "scipy (tested with 0.19)"); it is absent
from scipy 1.18 and, for three or more variables, it is a *randomised* lattice rule, so the reference's own
numbers carry ~1e-4 noise there.  This module restates the quantity with a fixed, documented rule so that
the oracle and the CUDA kernels can agree to round-off:

"SNQ" (shared-node quadrature).  Order the variables base-first, candidate-last.  With the Cholesky factor
``C_base = L L^T`` and ``z_base = m_base + L eta``, ``eta ~ N(0, I_t)``, the last variable is integrated
analytically,

    P(r_base, +) = E_eta[ 1{sign(z_base) = r_base} * Phi((m_c + l^T eta) / s) ],
    l = L^{-1} cov(base, c),   s^2 = var_c - |l|^2,

and the expectation over ``eta`` is a nested Gauss-Legendre rule applied *directly in eta*: dimension j is
split at the orthant boundary ``a_j(eta_<j) = -(m_j + sum_{i<j} L_ji eta_i) / L_jj`` into the two panels
``[-R, c]`` and ``[c, R]`` with ``c = clip(a_j, -R, R)``; the 2Q Gauss-Legendre nodes of the dimension are
shared out between the panels in proportion to their widths (``snq_split``; at least ``SNQ_QMIN`` each) and the
weights are multiplied by the standard normal density.  The node set depends only on the base variables, so one set
serves every candidate of a greedy step.  ``R = SNQ_R`` and ``Q = snq_order(t)`` for
t <= 5 base variables; from t = 6 on (batches of more than 6 samples) the tensor rule is replaced by a
sequential-conditioning lattice inside every base orthant (``sc_nodes``).

Nothing here is imported by the product path (ital_b200/); only tests/, bench.py's cpu_baseline /
``--impl reference`` legs and ``__graft_entry__.smoke()`` use it, as the checker.
"""

import numpy as np
from scipy.special import ndtr

SNQ_R = 7.0
SNQ_QMIN = 2
SNQ_WMIN = 1e-13      # nodes lighter than this are dropped (about half of them at t = 3, total mass ~1e-11)
_SQRT_2PI = np.sqrt(2.0 * np.pi)
_GL_CACHE = {}


SNQ_SC_FROM = 6        # bases of this many variables or more use the sequential-conditioning lattice
SNQ_SC_N = 262144      # ... with about this many nodes over all orthants
SNQ_SC_PILOT = 256     # nodes per orthant of the pilot pass that estimates the orthant masses
SNQ_SC_MIN = 64        # nodes per orthant at least
SNQ_SC_PMIN = 1e-13    # orthants lighter than this are left out
_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def snq_order(t):
    """Gauss-Legendre nodes per panel for a base of ``t`` <= 5 variables (0: sequential-conditioning lattice instead).

    Accuracy of the orthant probabilities / of the scores against converged rules (tests/test_orthant_vs_scipy.py,
    DESIGN.md): t = 1: 1e-11, t = 2: 1e-9, t = 3: 1e-7, t = 4: 1e-8 / 1e-7, t = 5: ~2e-6 / ~2e-5."""
    if t <= 1:
        return 32
    if t == 2:
        return 16
    if t in (3, 4):
        return 12
    if t == 5:
        return 10
    return 0


def sc_orthant(m_base, L_base, b, N):
    """N lattice nodes of N(m, L L^T) inside ONE orthant b (bit j set: variable j positive), whitened coordinates:
    (eta (N, t), w (N,)); sum(w) estimates the orthant probability.  See sc_nodes."""
    from scipy.special import ndtri
    m_base = np.asarray(m_base, dtype=np.float64)
    L_base = np.asarray(L_base, dtype=np.float64)
    t = len(m_base)
    alpha = np.array([np.sqrt(float(p)) - np.floor(np.sqrt(float(p))) for p in _PRIMES[:t]])
    k = np.arange(N, dtype=np.float64) + 0.5
    u = k[:, None] * alpha[None, :]
    u -= np.floor(u)
    u = 1.0 - np.abs(2.0 * u - 1.0)
    eta = np.zeros((N, t))
    w = np.full(N, 1.0 / N)
    for j in range(t):
        a = -(m_base[j] + eta[:, :j] @ L_base[j, :j]) / L_base[j, j]
        if (b >> j) & 1:
            q = ndtr(-a)
            eta[:, j] = -ndtri(np.maximum(u[:, j] * q, 1e-300))
        else:
            q = ndtr(a)
            eta[:, j] = ndtri(np.maximum(u[:, j] * q, 1e-300))
        w = w * q
    return eta, w


def sc_nodes(m_base, L_base, n_total=SNQ_SC_N):
    """Node set for t >= 6 base variables (batches of more than 6 samples), where a tensor rule explodes.

    For every orthant b of the base a Kronecker sequence u_kj = frac((k + 1/2) sqrt(p_j)), p_j the j-th prime, folded by
    the tent map 1 - |2u - 1|, is pushed through Genz's sequential conditioning INSIDE that orthant: eta_j is drawn
    from the standard normal truncated to the half-line on the orthant's side of the boundary
    a_j = -(m_j + sum_{i<j} L_ji eta_i) / L_jj and the node weight collects the half-line masses, so the integrand the
    candidates add (Phi of an affine function of eta) stays smooth on every node set.  A pilot pass of SNQ_SC_PILOT
    nodes per orthant estimates the orthant masses; ``n_total`` nodes are then shared out in proportion to them (at
    least SNQ_SC_MIN each; orthants below SNQ_SC_PMIN are left out) and the weights are scaled so that the orthant masses
    add up to one.  Nodes come out sorted by orthant.  Accuracy: 1e-4 class in the orthant probabilities
    (tests/test_orthant_vs_scipy.py) -- the reference's own ``mvndst(maxpts=100*dim, abseps=1e-4)``
    (ital/ital.py:380-381) is no better for these dimensions.
    """
    m_base = np.asarray(m_base, dtype=np.float64)
    L_base = np.asarray(L_base, dtype=np.float64)
    t = len(m_base)

    def gen(b, N):
        return sc_orthant(m_base, L_base, b, N)

    P = np.array([gen(b, SNQ_SC_PILOT)[1].sum() for b in range(1 << t)])
    etas, ws, orth = [], [], []
    for b in range(1 << t):
        if P[b] < SNQ_SC_PMIN:
            continue
        nb = max(SNQ_SC_MIN, int(np.floor(n_total * P[b] / P.sum() + 0.5)))
        e, w = gen(b, nb)
        etas.append(e)
        ws.append(w)
        orth.append(np.full(nb, b, dtype=np.int64))
    w = np.concatenate(ws)
    return np.concatenate(etas), w / w.sum(), np.concatenate(orth)      # (the orthant masses must add up to one)


def gauss_legendre(q):
    """Nodes/weights of the q-point Gauss-Legendre rule on [-1, 1] (numpy's Golub-Welsch + Newton)."""
    if q not in _GL_CACHE:
        _GL_CACHE[q] = np.polynomial.legendre.leggauss(q)
    return _GL_CACHE[q]


def _phi(x):
    return np.exp(-0.5 * x * x) / _SQRT_2PI


def snq_split(c, q, R=SNQ_R, q_min=SNQ_QMIN):
    """Number of the 2q nodes of one dimension given to the lower panel [-R, c] (rest: upper panel [c, R]).

    Proportional to the panel widths so that the node density is the same on both sides of the boundary.
    """
    n_lo = np.floor(2 * q * (c + R) / (2.0 * R) + 0.5).astype(np.int64)
    return np.clip(n_lo, q_min, 2 * q - q_min)


def snq_nodes(m_base, L_base, q=None, R=SNQ_R, w_min=SNQ_WMIN):
    """Shared nodes for a base N(m_base, L L^T).

    Returns ``eta`` (N, t), ``w`` (N,), ``orth`` (N,) with ``N <= (2q)^t`` (nodes whose weight is below ``w_min``
    are dropped after the full product rule has been formed).  Bit j of ``orth`` is 1 where
    ``z_j > 0``.  Node index = sum_j digit_j * (2q)^(t-1-j), digit_j in [0, 2q): the first n_lo digits
    of a dimension are the lower panel (z_j < 0), the rest the upper one; n_lo = snq_split(c).
    """
    m_base = np.asarray(m_base, dtype=np.float64)
    L_base = np.asarray(L_base, dtype=np.float64)
    t = len(m_base)
    if q is None:
        if t >= SNQ_SC_FROM:
            return sc_nodes(m_base, L_base)
        q = snq_order(t)
    eta = np.zeros((1, 0))
    w = np.ones(1)
    orth = np.zeros(1, dtype=np.int64)
    for j in range(t):
        a = -(m_base[j] + eta @ L_base[j, :j]) / L_base[j, j]
        c = np.clip(a, -R, R)
        n_lo = snq_split(c, q, R)
        N = len(c)
        x = np.empty((N, 2 * q))
        ww = np.empty((N, 2 * q))
        bit = np.empty((N, 2 * q), dtype=np.int64)
        for nl in np.unique(n_lo):
            sel = np.nonzero(n_lo == nl)[0]
            cs = c[sel]
            g, gw = gauss_legendre(int(nl))
            half = 0.5 * (cs + R)                    # lower panel [-R, c]
            xl = -R + half[:, None] * (1.0 + g)[None, :]
            x[sel, :nl] = xl
            ww[sel, :nl] = half[:, None] * gw[None, :] * _phi(xl)
            bit[sel, :nl] = 0
            g, gw = gauss_legendre(int(2 * q - nl))
            half = 0.5 * (R - cs)                    # upper panel [c, R]
            xh = cs[:, None] + half[:, None] * (1.0 + g)[None, :]
            x[sel, nl:] = xh
            ww[sel, nl:] = half[:, None] * gw[None, :] * _phi(xh)
            bit[sel, nl:] = 1
        eta = np.concatenate((np.repeat(eta, 2 * q, axis=0), x.reshape(-1, 1)), axis=1)
        w = (w[:, None] * ww).reshape(-1)
        orth = (orth[:, None] + (bit << j)).reshape(-1)
    if w_min > 0 and t > 0:
        keep = w >= w_min
        eta, w, orth = eta[keep], w[keep], orth[keep]
    return eta, w, orth


def base_masses(w, orth, t):
    """Quadrature estimate of the 2^t base orthant probabilities."""
    return np.bincount(orth, weights=w, minlength=1 << t)


def _joint_rows(args):
    """Worker of snq_joint: candidates [lo, hi) against the (sorted) shared nodes."""
    m_c, l_c, s_c, eta, w, bounds, chunk = args
    return _joint_block(m_c, l_c, s_c, eta, w, bounds, chunk)


def snq_joint(m_base, L_base, m_c, l_c, s_c, q=None, R=SNQ_R, chunk=None, pool=None, pool_tasks=64):
    """P(r_base, candidate > 0) and P(r_base) for many candidates sharing one base.

    m_c (n,), l_c (n, t), s_c (n,)  ->  p_plus (n, 2^t), p_base (2^t,).  ``s_c <= 0`` turns Phi into a step.
    ``pool``: a multiprocessing pool to spread the candidates over host cores -- the counterpart of the
    reference's ``parallelized=True`` ``Pool.map`` over candidates (ital/ital.py:124-126).
    """
    m_c = np.atleast_1d(np.asarray(m_c, dtype=np.float64))
    l_c = np.asarray(l_c, dtype=np.float64).reshape(len(m_c), -1)
    s_c = np.atleast_1d(np.asarray(s_c, dtype=np.float64))
    t = l_c.shape[1]
    eta, w, orth = snq_nodes(m_base, L_base, q, R)
    nb = 1 << t
    p_base = base_masses(w, orth, t)
    n = len(m_c)
    if chunk is None:
        chunk = max(1, int(4e6 // max(1, len(w))))
    order = np.argsort(orth, kind='stable')
    eta, w, orth = eta[order], w[order], orth[order]
    bounds = np.searchsorted(orth, np.arange(nb + 1))
    if pool is not None and n > 1:
        cuts = np.linspace(0, n, min(pool_tasks, n) + 1).astype(np.int64)
        parts = pool.map(_joint_rows, [(m_c[a:b], l_c[a:b], s_c[a:b], eta, w, bounds, chunk)
                                       for a, b in zip(cuts[:-1], cuts[1:]) if b > a])
        return np.concatenate(parts, axis=0), p_base
    return _joint_block(m_c, l_c, s_c, eta, w, bounds, chunk), p_base


def _joint_block(m_c, l_c, s_c, eta, w, bounds, chunk):
    n, nb = len(m_c), len(bounds) - 1
    p_plus = np.empty((n, nb))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        num = m_c[lo:hi, None] + l_c[lo:hi] @ eta.T
        with np.errstate(divide='ignore', invalid='ignore'):
            arg = num / s_c[lo:hi, None]
        deg = s_c[lo:hi] <= 0
        if np.any(deg):
            arg[deg] = np.where(num[deg] > 0, np.inf, -np.inf)
        cdf = ndtr(arg) * w[None, :]
        for b in range(nb):
            p_plus[lo:hi, b] = cdf[:, bounds[b]:bounds[b + 1]].sum(axis=1)
    return p_plus


def orthant_prob_all(mean, cov, q=None, R=SNQ_R):
    """All 2^D orthant probabilities of N(mean, cov) by SNQ with the last variable analytic.

    Returns an array ``p`` of length 2^D indexed by sum_j r_j << j (r_j = 1 for z_j > 0).
    D = 1 is the closed form of ital.py:364-369.
    """
    mean = np.asarray(mean, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64).reshape(len(mean), len(mean))
    D = len(mean)
    if D == 1:
        sd = np.sqrt(max(cov[0, 0], 0.0))
        if sd > 0:
            return np.array([ndtr(-mean[0] / sd), ndtr(mean[0] / sd)])
        return np.array([float(mean[0] <= 0), float(mean[0] > 0)])
    t = D - 1
    L = safe_cholesky(cov[:t, :t])
    l = np.linalg.solve(L, cov[:t, t]) if t > 0 else np.zeros(0)
    s2 = cov[t, t] - l @ l
    s = np.sqrt(s2) if s2 > 0 else 0.0
    p_plus, p_base = snq_joint(mean[:t], L, mean[t:t + 1], l[None, :], np.array([s]), q, R)
    out = np.empty(1 << D)
    out[(1 << t):] = p_plus[0]
    out[:(1 << t)] = p_base - p_plus[0]
    return out


def safe_cholesky(C, floor=1e-300):
    """Lower Cholesky factor with pivots floored (duplicated base points give a zero pivot)."""
    C = np.asarray(C, dtype=np.float64)
    t = C.shape[0]
    L = np.zeros((t, t))
    for j in range(t):
        d = C[j, j] - L[j, :j] @ L[j, :j]
        L[j, j] = np.sqrt(d) if d > floor else np.sqrt(floor)
        for i in range(j + 1, t):
            L[i, j] = (C[i, j] - L[i, :j] @ L[j, :j]) / L[j, j]
    return L


def orthant_prob(rel, mean, cov, q=None, R=SNQ_R):
    """P(sign(z) = rel) -- the quantity of MutualInformation.prob_rel (ital.py:345-383)."""
    rel = np.asarray(rel).astype(bool)
    idx = int(sum(int(r) << j for j, r in enumerate(rel)))
    return float(orthant_prob_all(mean, cov, q, R)[idx])
