"""Measured tolerance gate for a tensor-core Gram (north_star: "tcgen05 fed by TMA, in FP32-accurate split-TF32 or FP64
as the tolerance requires").

tcgen05 has no FP64 kind; the most accurate tensor-core route for the RBF Gram k(X, X_L) is the 3xTF32 split
(a = a_hi + a_lo with both parts in TF32, a.b ~ a_hi.b_hi + a_hi.b_lo + a_lo.b_hi) with FP32 accumulation in TMEM.
This script emulates exactly that arithmetic in numpy (TF32 = 10 explicit mantissa bits, round to nearest even;
products exact, accumulation in float32, K-chunks of 8 as the MMA accumulates), keeps everything downstream of the Gram
in float64 (as the CUDA path does), and measures the error of the posterior moments against the float64 path on the
data sets at hand.  The bar is the path's stated tolerance: 1e-6 relative on posterior means and variances.

    python tools/probe/tf32_gate.py            (CPU only; prints the table kept in profiles/r02_tf32_gate.txt)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def tf32(x):
    """Round float32 to TF32 (1 + 8 + 10 bits), nearest even."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def dot_3xtf32(A, B, chunk=8):
    """A (n, d) . B (m, d)^T with the 3xTF32 split and float32 accumulation."""
    A32, B32 = A.astype(np.float32), B.astype(np.float32)
    Ah, Bh = tf32(A32), tf32(B32)
    Al, Bl = tf32(A32 - Ah), tf32(B32 - Bh)
    acc = np.zeros((len(A), len(B)), dtype=np.float32)
    for k in range(0, A.shape[1], chunk):
        sl = slice(k, k + chunk)
        # products of TF32 operands are exact in float32-pairs; each MMA adds its chunk into the fp32 accumulator
        part = (Ah[:, sl].astype(np.float64) @ Bh[:, sl].T.astype(np.float64)
                + Ah[:, sl].astype(np.float64) @ Bl[:, sl].T.astype(np.float64)
                + Al[:, sl].astype(np.float64) @ Bh[:, sl].T.astype(np.float64))
        acc = (acc.astype(np.float64) + part).astype(np.float32)
    return acc.astype(np.float64)


def moments(X, lab, y, ls, gram):
    """Posterior mean / variance of every row given the labelled set, float64 downstream of the row.X_L products."""
    XL = X[lab]
    sq = np.sum(X ** 2, axis=1)
    dots = gram(X, XL)
    K_nL = np.exp((sq[:, None] + sq[lab][None, :] - 2.0 * dots) / (-2.0 * ls * ls))
    dLL = XL @ XL.T
    K = np.exp((sq[lab][:, None] + sq[lab][None, :] - 2.0 * dLL) / (-2.0 * ls * ls)) + 1e-6 * np.eye(len(lab))
    Ki = np.linalg.inv(K)
    m = K_nL @ (Ki @ y)
    v = 1.0 - np.sum((K_nL @ Ki) * K_nL, axis=1)
    return m, v, np.linalg.cond(K)


def main():
    rng = np.random.default_rng(0)
    sets = []
    ref = '/root/reference/data/butterflies_pca50.npz'
    if os.path.exists(ref):
        d = np.load(ref)
        Xb = d['X_train']
        sets.append(('butterflies 1000x50 (ls 2.5)', (Xb - Xb.min()) / (Xb.max() - Xb.min()), 2.5))
    import bench
    Xs, _ = bench.syn_block(0, 4000, 512)
    sets.append(('SYN 4000x512 float32 (ls 1.0)', Xs.astype(np.float64), 1.0))
    print('%-32s %9s  %-28s %-28s' % ('data', 'cond(K)', '3xTF32 + fp32 accumulate', 'float32 in, float64 accumulate'))
    for name, X, ls in sets:
        lab = rng.choice(len(X), 41, replace=False)
        y = rng.choice([-1.0, 1.0], 41)
        m0, v0, cond = moments(X, lab, y, ls, lambda A, B: A @ B.T)
        out = []
        for gram in (dot_3xtf32, lambda A, B: A.astype(np.float32).astype(np.float64) @ B.astype(np.float32).astype(np.float64).T):
            m1, v1, _ = moments(X, lab, y, ls, gram)
            dm, dv = np.max(np.abs(m1 - m0)), np.max(np.abs(v1 - v0))          # absolute (prior variance 1, |mean| <= ~1)
            out.append('dm %.1e  dv %.1e  %s' % (dm, dv, 'FAILS 1e-6' if max(dm, dv) > 1e-6 else 'passes'))
        print('%-32s %9.1e  %-28s %-28s' % (name, cond, out[0], out[1]))


if __name__ == '__main__':
    main()
