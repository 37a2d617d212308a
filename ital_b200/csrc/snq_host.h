// Shared-node quadrature ("SNQ") for Gaussian orthant probabilities: host-side node generation.
//
// Replaces the per-candidate, per-configuration calls of MutualInformation.prob_rel into
// scipy.stats.mvn.mvndst (/root/reference/ital/ital.py:373-383): all candidates of one greedy step share the
// base variables, so one node set in the whitened base coordinates serves every candidate.
// The rule is specified in DESIGN.md ("SNQ") and restated independently by oracle/orthant.py:
//   dimension j is split at c = clip(a_j, -R, R), a_j = -(m_j + sum_{i<j} L_ji eta_i) / L_jj, into the panels
//   [-R, c] and [c, R]; the 2q Gauss-Legendre nodes of the dimension are shared out between the two panels in
//   proportion to their widths (at least SNQ_QMIN each); weights carry the standard normal density.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace snq {

// Host node generation for long batches is millions of independent nodes: spread index ranges over the host cores.
template <typename F>
inline void parallel_for(int64_t n, int64_t grain, F&& body) {
    const int64_t max_threads = std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), 32));
    const int64_t threads = std::max<int64_t>(1, std::min<int64_t>(max_threads, n / std::max<int64_t>(1, grain)));
    if (threads <= 1) {
        body((int64_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int64_t th = 0; th < threads; ++th) {
        const int64_t lo = th * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        pool.emplace_back([&body, lo, hi]() { body(lo, hi); });
    }
    for (auto& th : pool) th.join();
}

constexpr double kR = 7.0;
constexpr int kQMin = 2;
constexpr int kMaxOrder = 64;
constexpr double kWMin = 1e-13;   // nodes lighter than this are dropped (about half of them at t = 3, mass ~1e-11)

constexpr int kScFrom = 6;         // bases of this many variables or more: sequential-conditioning lattice instead
constexpr int64_t kScN = 262144;   // ... with about this many nodes over all orthants
constexpr int kScPilot = 256;      // nodes per orthant of the pilot pass that estimates the orthant masses
constexpr int kScMin = 64;         // nodes per orthant at least
constexpr double kScPMin = 1e-13;  // orthants lighter than this are left out

inline int order_for(int t) {      // Gauss-Legendre nodes per panel (0: sequential-conditioning lattice instead)
    if (t <= 1) return 32;
    if (t == 2) return 16;
    if (t == 3 || t == 4) return 12;
    if (t == 5) return 10;
    return 0;
}

inline int64_t capacity_for(int t) {
    if (t >= kScFrom) return kScN + ((int64_t)kScMin << t);
    int64_t n = 1;
    for (int j = 0; j < t; ++j) n *= 2 * order_for(t);
    return n;
}

// inverse of the standard normal CDF: Abramowitz-Stegun 26.2.23 start, Halley steps on 0.5 erfc(-x / sqrt 2)
inline double ndtri(double p) {
    const bool lower = p < 0.5;
    const double pp = lower ? p : 1.0 - p;
    const double tt = std::sqrt(-2.0 * std::log(pp));
    double x = tt - (2.515517 + 0.802853 * tt + 0.010328 * tt * tt) /
                        (1.0 + 1.432788 * tt + 0.189269 * tt * tt + 0.001308 * tt * tt * tt);
    x = lower ? -x : x;
    for (int it = 0; it < 6; ++it) {
        const double cdf = 0.5 * std::erfc(-x * 0.70710678118654752440);
        const double pdf = std::exp(-0.5 * x * x) * 0.39894228040143267794;
        const double f = cdf - p;
        const double dx = f / (pdf + 0.5 * x * f);        // Halley: f / (f' - f f'' / (2 f')), f'' = -x f'
        x -= dx;
        if (std::fabs(dx) < 1e-15 * (1.0 + std::fabs(x))) break;
    }
    return x;
}

struct GaussLegendre {
    std::vector<double> x[kMaxOrder + 1], w[kMaxOrder + 1];
    GaussLegendre() {
        const double pi = 3.14159265358979323846;
        for (int n = 1; n <= kMaxOrder; ++n) {
            x[n].resize(n);
            w[n].resize(n);
            for (int i = 0; i < n; ++i) {
                // i-th root counted from the right; Newton on P_n with the classical cosine start.
                double z = std::cos(pi * (i + 0.75) / (n + 0.5));
                double pp = 0.0;
                for (int it = 0; it < 100; ++it) {
                    double p1 = 1.0, p2 = 0.0;
                    for (int k = 1; k <= n; ++k) {
                        double p3 = p2;
                        p2 = p1;
                        p1 = ((2.0 * k - 1.0) * z * p2 - (k - 1.0) * p3) / k;
                    }
                    pp = n * (z * p1 - p2) / (z * z - 1.0);
                    double dz = p1 / pp;
                    z -= dz;
                    if (std::fabs(dz) < 1e-16) break;
                }
                // recompute the derivative at the converged root for the weight
                double p1 = 1.0, p2 = 0.0;
                for (int k = 1; k <= n; ++k) {
                    double p3 = p2;
                    p2 = p1;
                    p1 = ((2.0 * k - 1.0) * z * p2 - (k - 1.0) * p3) / k;
                }
                pp = n * (z * p1 - p2) / (z * z - 1.0);
                x[n][n - 1 - i] = z;                       // ascending order
                w[n][n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
            }
        }
    }
};

inline const GaussLegendre& gl() {
    static const GaussLegendre table;
    return table;
}

inline double phi(double x) { return std::exp(-0.5 * x * x) / std::sqrt(2.0 * 3.14159265358979323846); }

struct Nodes {
    int t = 0;
    int64_t n = 0;
    std::vector<double> eta;          // dimension-major: eta[j * n + k]
    std::vector<double> w;
    std::vector<int32_t> orth;        // bit j set where base variable j is positive
    std::vector<int32_t> group_begin; // 2^t + 1 offsets after sorting by orthant
    std::vector<double> masses;       // quadrature estimate of the 2^t base orthant probabilities
    double entropy = 0.0;             // score of the base alone: sum P (log(1+eps) - log(P+eps))
};

// Node set for t >= kScFrom base variables (batches of more than 6 samples), where a tensor rule explodes: for every
// orthant b of the base a Kronecker sequence u_k = frac((k + 1/2) sqrt(p_j)) (p_j the j-th prime), folded by the
// tent map 1 - |2u - 1|, is pushed through Genz's sequential conditioning INSIDE that orthant -- eta_j is drawn from
// the standard normal truncated to the half-line on the orthant's side of the boundary
// a_j = -(m_j + sum_{i<j} L_ji eta_i) / L_jj, the node weight collects the half-line masses -- so that the integrand
// the candidates add (Phi of an affine function of eta) stays smooth on every node set.  A pilot pass of kScPilot
// nodes per orthant estimates the orthant masses; the kScN nodes are then shared out in proportion to them (at least
// kScMin each; orthants below kScPMin are left out) and the weights scaled so that the orthant masses add up to one.
// Nodes come out sorted by orthant.  Accuracy: 1e-4 class in the orthant probabilities (tests/test_orthant_vs_scipy.py).
// Nodes of ONE orthant b (bit j set: z_j > 0): eta dimension-major with stride `stride` (nullptr: weights only).
inline void sc_orthant(int t, const double* m, const double* L, int b, int64_t N, double* eta, int64_t stride, double* w) {
    static const int primes[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    double alpha[12];
    for (int j = 0; j < t; ++j) {
        alpha[j] = std::sqrt((double)primes[j]);
        alpha[j] -= std::floor(alpha[j]);
    }
    auto cdf = [](double x) { return 0.5 * std::erfc(-x * 0.70710678118654752440); };
    std::vector<double> e(t);
    for (int64_t k = 0; k < N; ++k) {
        double wk = 1.0 / (double)N;
        for (int j = 0; j < t; ++j) {
            double u = ((double)k + 0.5) * alpha[j];
            u -= std::floor(u);
            u = 1.0 - std::fabs(2.0 * u - 1.0);
            double acc = m[j];
            for (int i = 0; i < j; ++i) acc += L[j * t + i] * e[i];
            const double a = -acc / L[j * t + j];
            if ((b >> j) & 1) {                     // z_j > 0: eta_j above a
                const double q = cdf(-a);
                const double p = u * q;
                e[j] = -ndtri(p < 1e-300 ? 1e-300 : p);
                wk *= q;
            } else {
                const double q = cdf(a);
                const double p = u * q;
                e[j] = ndtri(p < 1e-300 ? 1e-300 : p);
                wk *= q;
            }
            if (eta) eta[(size_t)j * stride + k] = e[j];
        }
        w[k] = wk;
    }
}

inline Nodes generate_sc(int t, const double* m, const double* L) {
    const int nb = 1 << t;
    // nodes of orthant b: eta (dimension-major, stride N) and weights
    auto gen = [&](int b, int64_t N, double* eta, int64_t stride, double* w) { sc_orthant(t, m, L, b, N, eta, stride, w); };
    // (orthants are independent: the pilot pass and the final pass run over the host cores, one orthant at a time per
    // thread; masses are summed per orthant in node order, so the result does not depend on the thread count)
    std::vector<double> P(nb);
    parallel_for(nb, 1, [&](int64_t b_lo, int64_t b_hi) {
        std::vector<double> wp(kScPilot);
        for (int64_t b = b_lo; b < b_hi; ++b) {
            gen((int)b, kScPilot, nullptr, 0, wp.data());
            double s = 0.0;
            for (int k = 0; k < kScPilot; ++k) s += wp[k];
            P[b] = s;
        }
    });
    double total = 0.0;
    for (int b = 0; b < nb; ++b) total += P[b];
    std::vector<int64_t> cnt(nb, 0), start(nb + 1, 0);
    for (int b = 0; b < nb; ++b) {
        if (P[b] >= kScPMin) cnt[b] = std::max<int64_t>(kScMin, (int64_t)std::floor((double)kScN * P[b] / total + 0.5));
        start[b + 1] = start[b] + cnt[b];
    }
    const int64_t n = start[nb];
    Nodes out;
    out.t = t;
    out.n = n;
    out.eta.assign((size_t)t * n, 0.0);
    out.w.assign(n, 0.0);
    out.orth.assign(n, 0);
    out.group_begin.assign(nb + 1, 0);
    out.masses.assign(nb, 0.0);
    for (int b = 0; b <= nb; ++b) out.group_begin[b] = (int32_t)start[b];
    parallel_for(nb, 1, [&](int64_t b_lo, int64_t b_hi) {
        for (int64_t b = b_lo; b < b_hi; ++b) {
            if (cnt[b] == 0) continue;
            const int64_t pos = start[b];
            gen((int)b, cnt[b], out.eta.data() + pos, n, out.w.data() + pos);
            double s = 0.0;
            for (int64_t k = 0; k < cnt[b]; ++k) {
                s += out.w[pos + k];
                out.orth[pos + k] = (int32_t)b;
            }
            out.masses[b] = s;
        }
    });
    // the orthant masses must add up to one: scaling the weights accordingly removes the error all node sets share
    double tot = 0.0;
    for (int b = 0; b < nb; ++b) tot += out.masses[b];
    for (double& wk : out.w) wk /= tot;
    for (int b = 0; b < nb; ++b) out.masses[b] /= tot;
    out.entropy = 0.0;
    const double eps = 1e-12, log1p_eps = std::log(1.0 + eps);
    for (int b = 0; b < nb; ++b) out.entropy += out.masses[b] * (log1p_eps - std::log(out.masses[b] + eps));
    return out;
}

// m[t], L[t*t] row-major lower triangular.  Node k = sum_j g_j (2q)^(t-1-j) (the digit of the first dimension varies
// slowest); contiguous index ranges are generated independently over the host cores, each walking its range like an
// odometer (a dimension is recomputed only when its digit or one before it changes), the kept nodes (weight >= kWMin)
// are then scattered to their place in the orthant-sorted list, in index order within an orthant.
inline Nodes generate(int t, const double* m, const double* L, int q = 0, double R = kR) {
    if (q <= 0 && t >= kScFrom) return generate_sc(t, m, L);
    Nodes out;
    out.t = t;
    if (q <= 0) q = order_for(t);
    const int two_q = 2 * q;
    const int nb = 1 << t;
    int64_t N = 1;
    for (int j = 0; j < t; ++j) N *= two_q;
    const GaussLegendre& G = gl();
    // chunks of consecutive indices; per chunk the kept nodes in generation order
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(1024, N / 2048));
    struct Chunk {
        std::vector<double> eta;      // node-major, t per node
        std::vector<double> w;
        std::vector<int32_t> orth;
        std::vector<int32_t> cnt;     // kept per orthant
    };
    std::vector<Chunk> chunks(n_chunks);
    parallel_for(n_chunks, 1, [&](int64_t c_lo, int64_t c_hi) {
        for (int64_t c = c_lo; c < c_hi; ++c) {
            Chunk& ch = chunks[c];
            ch.cnt.assign(nb, 0);
            const int64_t k_lo = N * c / n_chunks, k_hi = N * (c + 1) / n_chunks;
            int dig[16], bit[16];
            double e[16], wpre[17];
            wpre[0] = 1.0;
            int from = 0;                                // first dimension whose state is stale
            for (int64_t k = k_lo; k < k_hi; ++k) {
                if (k == k_lo) {
                    int64_t rem = k;
                    for (int j = t - 1; j >= 0; --j) { dig[j] = (int)(rem % two_q); rem /= two_q; }
                    from = 0;
                }
                for (int j = from; j < t; ++j) {
                    double acc = m[j];
                    for (int i = 0; i < j; ++i) acc += e[i] * L[j * t + i];
                    const double a = -acc / L[j * t + j];
                    const double cc = a < -R ? -R : (a > R ? R : a);
                    int n_lo = (int)std::floor(two_q * (cc + R) / (2.0 * R) + 0.5);
                    if (n_lo < kQMin) n_lo = kQMin;
                    if (n_lo > two_q - kQMin) n_lo = two_q - kQMin;
                    const int n_hi = two_q - n_lo;
                    const double half_lo = 0.5 * (cc + R), half_hi = 0.5 * (R - cc);
                    const int g = dig[j];
                    double xg, wg;
                    if (g < n_lo) {
                        xg = -R + half_lo * (1.0 + G.x[n_lo][g]);
                        wg = half_lo * G.w[n_lo][g] * phi(xg);
                        bit[j] = 0;
                    } else {
                        xg = cc + half_hi * (1.0 + G.x[n_hi][g - n_lo]);
                        wg = half_hi * G.w[n_hi][g - n_lo] * phi(xg);
                        bit[j] = 1;
                    }
                    e[j] = xg;
                    wpre[j + 1] = wpre[j] * wg;
                }
                const double wk = wpre[t];
                if (t == 0 || wk >= kWMin) {
                    int ob = 0;
                    for (int j = 0; j < t; ++j) ob |= bit[j] << j;
                    for (int j = 0; j < t; ++j) ch.eta.push_back(e[j]);
                    ch.w.push_back(wk);
                    ch.orth.push_back(ob);
                    ch.cnt[ob]++;
                }
                // odometer: advance the digits, remember the first dimension that changed
                int j = t - 1;
                while (j >= 0 && dig[j] == two_q - 1) { dig[j] = 0; --j; }
                if (j >= 0) dig[j]++;
                from = j < 0 ? 0 : j;
            }
        }
    });
    // offsets: orthant-major, chunk-minor
    out.group_begin.assign(nb + 1, 0);
    std::vector<int64_t> pos((size_t)n_chunks * nb);
    int64_t total = 0;
    for (int b = 0; b < nb; ++b) {
        out.group_begin[b] = (int32_t)total;
        for (int64_t c = 0; c < n_chunks; ++c) {
            pos[(size_t)c * nb + b] = total;
            total += chunks[c].cnt[b];
        }
    }
    out.group_begin[nb] = (int32_t)total;
    const int64_t n = total;
    out.n = n;
    out.eta.assign((size_t)t * n, 0.0);
    out.w.assign(n, 0.0);
    out.orth.assign(n, 0);
    parallel_for(n_chunks, 1, [&](int64_t c_lo, int64_t c_hi) {
        for (int64_t c = c_lo; c < c_hi; ++c) {
            const Chunk& ch = chunks[c];
            std::vector<int64_t> cur(pos.begin() + (size_t)c * nb, pos.begin() + (size_t)(c + 1) * nb);
            for (size_t k = 0; k < ch.w.size(); ++k) {
                const int64_t p = cur[ch.orth[k]]++;
                for (int i = 0; i < t; ++i) out.eta[(size_t)i * n + p] = ch.eta[k * t + i];
                out.w[p] = ch.w[k];
                out.orth[p] = ch.orth[k];
            }
        }
    });
    out.masses.assign(nb, 0.0);
    for (int b = 0; b < nb; ++b) {
        double sum = 0.0;
        for (int32_t k = out.group_begin[b]; k < out.group_begin[b + 1]; ++k) sum += out.w[k];
        out.masses[b] = sum;
    }
    // same summand as the candidate scores (ital.py:207-219 with p' = 1), so that gain = score - entropy
    out.entropy = 0.0;
    const double eps = 1e-12, log1p_eps = std::log(1.0 + eps);
    for (int b = 0; b < nb; ++b) out.entropy += out.masses[b] * (log1p_eps - std::log(out.masses[b] + eps));
    return out;
}

}  // namespace snq

// ---------------------------------------------------------------------------------------------------------------
// Node sets of the GENERAL feedback model (label_prob < 1): the user may skip samples, so the score of a candidate
// needs, besides P(r), the orthant probabilities of the unobserved samples conditional on pinned (labelled) ones
// (MutualInformation.updated_prob_rel, /root/reference/ital/ital.py:432-450, via gp.updated_prediction,
// /root/reference/ital/gp.py:295-344).  For every subset O of the batch's base that is labelled and every sign
// pattern f of those labels, the whitened coordinates eta of the base are Gaussian with
//     mu = A^T (A A^T + noise I)^-1 (f - m_O),   Sigma = I - A^T (A A^T + noise I)^-1 A,   A = L[O, :],
// and the unlabelled base variables z_U = m_U + L[U, :] eta get their own SNQ nodes zeta; eta = mu + G zeta with
// G = Sigma L[U,:]^T Lc^-T, Lc Lc^T = L[U,:] Sigma L[U,:]^T.  Every set is a list of (eta, weight, orthant of z_U)
// shared by all candidates, exactly like the unconditional nodes (set 0).  See DESIGN.md "general feedback model"
// and tools/proto_general.py (numpy prototype checked against the oracle and the reference goldens).
namespace snq {

struct GeneralSets {
    int t = 0;
    int n_sets = 0;
    int n_groups = 0;                  // total number of (set, orthant of U) accumulators
    int64_t n_nodes = 0;
    std::vector<double> eta;           // dimension-major [t][n_nodes]
    std::vector<double> w;
    std::vector<int32_t> group_begin;  // n_groups + 1 node offsets (nodes sorted by global group)
    std::vector<double> group_mass;    // n_groups: sum of the weights of the group
    std::vector<int32_t> set_group0;   // n_sets + 1: first global group of each set
    // lut[r * 2^D + Omask] for r in [0, 2^D), Omask in [1, 2^D): {global group, set, flags}; D = t + 1, the
    // candidate is variable t.  flags: 1 = candidate labelled (use the density sums), 2 = every sample labelled.
    std::vector<int32_t> lut;
};

inline void chol_inplace(std::vector<double>& C, int n, double floor = 1e-300) {
    for (int j = 0; j < n; ++j) {
        double d = C[j * n + j];
        for (int k = 0; k < j; ++k) d -= C[j * n + k] * C[j * n + k];
        const double p = std::sqrt(d > floor ? d : floor);
        C[j * n + j] = p;
        for (int i = j + 1; i < n; ++i) {
            double v = C[i * n + j];
            for (int k = 0; k < j; ++k) v -= C[i * n + k] * C[j * n + k];
            C[i * n + j] = v / p;
        }
        for (int i = 0; i < j; ++i) C[i * n + j] = 0.0;
    }
}

// m[t], L[t*t] row-major lower triangular, label noise sigma^2
inline GeneralSets generate_general(int t, const double* m, const double* L, double noise) {
    GeneralSets out;
    out.t = t;
    const int D = t + 1;
    std::vector<std::vector<double>> eta_sets, w_sets;
    std::vector<std::vector<int32_t>> gb_sets;
    std::vector<int> set_Omask, set_fbits, set_nU;
    // set index by (Omask over base, sign bits compacted over O)
    std::vector<std::vector<int>> set_of(1 << t);
    for (int Omask = 0; Omask < (1 << t); ++Omask) {
        std::vector<int> O, U;
        for (int j = 0; j < t; ++j) ((Omask >> j) & 1 ? O : U).push_back(j);
        const int k = (int)O.size(), u = (int)U.size();
        set_of[Omask].assign(1 << k, -1);
        // pieces that do not depend on the signs
        std::vector<double> A((size_t)k * t), X((size_t)k * t, 0.0), Sig((size_t)t * t, 0.0);
        for (int a = 0; a < k; ++a)
            for (int c = 0; c < t; ++c) A[a * t + c] = L[O[a] * t + c];
        std::vector<double> S((size_t)k * k);
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) {
                double acc = (a == b) ? noise : 0.0;
                for (int c = 0; c < t; ++c) acc += A[a * t + c] * A[b * t + c];
                S[a * k + b] = acc;
            }
        if (k > 0) {
            chol_inplace(S, k);
            // X = S^-1 A by two triangular solves per column
            for (int c = 0; c < t; ++c) {
                std::vector<double> y(k);
                for (int a = 0; a < k; ++a) {
                    double v = A[a * t + c];
                    for (int b = 0; b < a; ++b) v -= S[a * k + b] * y[b];
                    y[a] = v / S[a * k + a];
                }
                for (int a = k - 1; a >= 0; --a) {
                    double v = y[a];
                    for (int b = a + 1; b < k; ++b) v -= S[b * k + a] * X[b * t + c];
                    X[a * t + c] = v / S[a * k + a];
                }
            }
        }
        for (int r = 0; r < t; ++r)
            for (int c = 0; c < t; ++c) {
                double acc = (r == c) ? 1.0 : 0.0;
                for (int a = 0; a < k; ++a) acc -= A[a * t + r] * X[a * t + c];
                Sig[r * t + c] = acc;
            }
        // B = L[U,:], CU = B Sig B^T, Lc = chol(CU), G = Sig B^T Lc^-T
        std::vector<double> B((size_t)u * t), SB((size_t)t * u), CU((size_t)u * u), G((size_t)t * u);
        for (int a = 0; a < u; ++a)
            for (int c = 0; c < t; ++c) B[a * t + c] = L[U[a] * t + c];
        for (int r = 0; r < t; ++r)
            for (int a = 0; a < u; ++a) {
                double acc = 0.0;
                for (int c = 0; c < t; ++c) acc += Sig[r * t + c] * B[a * t + c];
                SB[r * u + a] = acc;
            }
        for (int a = 0; a < u; ++a)
            for (int b = 0; b < u; ++b) {
                double acc = 0.0;
                for (int c = 0; c < t; ++c) acc += B[a * t + c] * SB[c * u + b];
                CU[a * u + b] = acc;
            }
        std::vector<double> Lc(CU);
        if (u > 0) chol_inplace(Lc, u);
        for (int r = 0; r < t; ++r)       // row r of G solves Lc x = (SB row r)
            for (int a = 0; a < u; ++a) {
                double v = SB[r * u + a];
                for (int b = 0; b < a; ++b) v -= Lc[a * u + b] * G[r * u + b];
                G[r * u + a] = v / Lc[a * u + a];
            }
        for (int fb = 0; fb < (1 << k); ++fb) {
            std::vector<double> mu(t, 0.0), resid(k);
            for (int a = 0; a < k; ++a) resid[a] = (((fb >> a) & 1) ? 1.0 : -1.0) - m[O[a]];
            for (int c = 0; c < t; ++c)
                for (int a = 0; a < k; ++a) mu[c] += X[a * t + c] * resid[a];
            std::vector<double> eta, w;
            std::vector<int32_t> gb;
            if (u == 0) {
                eta = mu;                  // a single node
                w.assign(1, 1.0);
                gb = {0, 1};
            } else {
                std::vector<double> mU(u);
                for (int a = 0; a < u; ++a) {
                    double acc = m[U[a]];
                    for (int c = 0; c < t; ++c) acc += B[a * t + c] * mu[c];
                    mU[a] = acc;
                }
                Nodes nd = generate(u, mU.data(), Lc.data());
                eta.assign((size_t)t * nd.n, 0.0);
                for (int64_t q = 0; q < nd.n; ++q)
                    for (int r = 0; r < t; ++r) {
                        double acc = mu[r];
                        for (int a = 0; a < u; ++a) acc += G[r * u + a] * nd.eta[(size_t)a * nd.n + q];
                        eta[(size_t)r * nd.n + q] = acc;
                    }
                w = nd.w;
                gb = nd.group_begin;
            }
            set_of[Omask][fb] = (int)eta_sets.size();
            eta_sets.push_back(eta);
            w_sets.push_back(w);
            gb_sets.push_back(gb);
            set_Omask.push_back(Omask);
            set_fbits.push_back(fb);
            set_nU.push_back(u);
        }
    }
    // flatten
    out.n_sets = (int)eta_sets.size();
    out.set_group0.assign(out.n_sets + 1, 0);
    int64_t total = 0;
    for (int s = 0; s < out.n_sets; ++s) {
        out.set_group0[s + 1] = out.set_group0[s] + (1 << set_nU[s]);
        total += (int64_t)w_sets[s].size();
    }
    out.n_groups = out.set_group0[out.n_sets];
    out.n_nodes = total;
    out.eta.assign((size_t)t * total, 0.0);
    out.w.resize(total);
    out.group_begin.assign(out.n_groups + 1, 0);
    out.group_mass.assign(out.n_groups, 0.0);
    int64_t pos = 0;
    for (int s = 0; s < out.n_sets; ++s) {
        const int64_t ns = (int64_t)w_sets[s].size();
        for (int r = 0; r < t; ++r)
            for (int64_t q = 0; q < ns; ++q) out.eta[(size_t)r * total + pos + q] = eta_sets[s][(size_t)r * ns + q];
        for (int64_t q = 0; q < ns; ++q) out.w[pos + q] = w_sets[s][q];
        for (int g = 0; g < (1 << set_nU[s]); ++g) {
            const int gg = out.set_group0[s] + g;
            out.group_begin[gg] = (int32_t)(pos + gb_sets[s][g]);
            double acc = 0.0;
            for (int32_t q = gb_sets[s][g]; q < gb_sets[s][g + 1]; ++q) acc += w_sets[s][q];
            out.group_mass[gg] = acc;
        }
        pos += ns;
    }
    out.group_begin[out.n_groups] = (int32_t)total;
    // lookup: (relevance configuration r of base + candidate, labelled subset O) -> accumulator
    out.lut.assign((size_t)3 << (2 * D), 0);
    for (int r = 0; r < (1 << D); ++r)
        for (int Om = 1; Om < (1 << D); ++Om) {
            const int Ob = Om & ((1 << t) - 1);
            const bool c_in = (Om >> t) & 1;
            int fb = 0, kk = 0, g = 0, uu = 0;
            for (int j = 0; j < t; ++j) {
                if ((Ob >> j) & 1) { fb |= ((r >> j) & 1) << kk; ++kk; }
                else { g |= ((r >> j) & 1) << uu; ++uu; }
            }
            const int s = set_of[Ob][fb];
            int32_t* e = &out.lut[((size_t)r * (1 << D) + Om) * 3];
            e[0] = out.set_group0[s] + g;
            e[1] = s;
            e[2] = (c_in ? 1 : 0) | (Om == (1 << D) - 1 ? 2 : 0);
        }
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// Node sets of the change-estimation subset (MutualInformation._call_iter_sub, /root/reference/ital/ital.py:227-275):
// the variables are ext = [B (tB samples of the batch), S' (subset members outside the batch)], D of them, with means
// m and Cholesky factor L (row-major D x D) of their posterior covariance.  Two families of 2^tB groups each, in the
// whitened coordinates eta of ext (consumed by k_eval_sub, which documents what each part is for):
//   part 1: the unconditional nodes of B alone (coordinates of S' zero);
//   part 2: nodes of the prior inside the orthant (r_B, s*), s* = signs of the means of S';
// and the moments of S' conditional on the labels r_B of B (labels +-1 with noise sigma^2) from which the kernel
// conditions on every candidate's own label.
// Orthants of up to 5 variables are cut out of the tensor rule, larger ones get kSubN lattice nodes each.
constexpr int64_t kSubN = 16384;

struct SubSets {
    int tB = 0, D = 0, n_groups = 0, sub_bits = 0;
    int64_t n_nodes = 0;
    std::vector<double> eta;            // dimension-major [D][n_nodes]
    std::vector<double> w;
    std::vector<int32_t> group_begin;   // 2 * 2^tB + 1
    std::vector<double> mass;           // [2][2^tB]: group masses of parts 1 and 2
    std::vector<double> mu;             // [2^tB][D]: mean of eta given the labels r_B
    std::vector<double> Sig;            // [D][D]: covariance of eta given labels on B
    std::vector<double> mU;             // [2^tB][u]: mean of S' given the labels r_B
    std::vector<double> CU;             // [u][u]: covariance of S' given labels on B
    std::vector<double> BS;             // [u][D]: Bm Sig, covariance of S' with eta given labels on B
};

// nodes of N(m, L L^T) (u variables) inside orthant b, coordinates whitened by L: appended to eta (node-major, u per
// node) and w
inline void orthant_nodes(int u, const double* m, const double* L, int b, std::vector<double>& eta, std::vector<double>& w) {
    if (u == 0) {
        w.push_back(1.0);
        return;
    }
    if (u < kScFrom) {
        Nodes nd = generate(u, m, L);
        for (int32_t q = nd.group_begin[b]; q < nd.group_begin[b + 1]; ++q) {
            for (int a = 0; a < u; ++a) eta.push_back(nd.eta[(size_t)a * nd.n + q]);
            w.push_back(nd.w[q]);
        }
        return;
    }
    std::vector<double> e((size_t)u * kSubN), ww(kSubN);
    // (one orthant: spread the nodes over the cores by running the lattice in slices would change nothing in the
    // result, but 16384 nodes take about a millisecond)
    sc_orthant(u, m, L, b, kSubN, e.data(), kSubN, ww.data());
    for (int64_t q = 0; q < kSubN; ++q) {
        for (int a = 0; a < u; ++a) eta.push_back(e[(size_t)a * kSubN + q]);
        w.push_back(ww[q]);
    }
}

inline SubSets generate_sub(int tB, int D, const double* m, const double* L, double noise) {
    SubSets out;
    out.tB = tB;
    out.D = D;
    const int u = D - tB, G = 1 << tB;
    out.n_groups = 2 * G;
    for (int a = 0; a < u; ++a)
        if (m[tB + a] > 0.0) out.sub_bits |= 1 << a;
    std::vector<std::vector<double>> g_eta(2 * G), g_w(2 * G);      // node-major, D coordinates per node
    out.mass.assign(2 * G, 0.0);
    // part 1: B alone
    if (tB == 0) {
        g_eta[0].assign(D, 0.0);
        g_w[0].assign(1, 1.0);
        out.mass[0] = 1.0;
    } else {
        std::vector<double> LB((size_t)tB * tB);
        for (int a = 0; a < tB; ++a)
            for (int c = 0; c < tB; ++c) LB[a * tB + c] = L[a * D + c];
        Nodes nd = generate(tB, m, LB.data());
        for (int g = 0; g < G; ++g) {
            for (int32_t q = nd.group_begin[g]; q < nd.group_begin[g + 1]; ++q) {
                for (int c = 0; c < D; ++c) g_eta[g].push_back(c < tB ? nd.eta[(size_t)c * nd.n + q] : 0.0);
                g_w[g].push_back(nd.w[q]);
            }
            out.mass[g] = nd.masses[g];
        }
    }
    // the pieces of the conditioning on B that do not depend on the labels: X = (A A^T + noise I)^-1 A, A = L[B, :],
    // Sig = I - A^T X; S' = m_U + Bm eta, CU = Bm Sig Bm^T
    std::vector<double> X((size_t)tB * D, 0.0), Sig((size_t)D * D, 0.0);
    if (tB > 0) {
        std::vector<double> S((size_t)tB * tB);
        for (int a = 0; a < tB; ++a)
            for (int b = 0; b < tB; ++b) {
                double acc = (a == b) ? noise : 0.0;
                for (int c = 0; c < D; ++c) acc += L[a * D + c] * L[b * D + c];
                S[a * tB + b] = acc;
            }
        chol_inplace(S, tB);
        for (int c = 0; c < D; ++c) {
            std::vector<double> y(tB);
            for (int a = 0; a < tB; ++a) {
                double v = L[a * D + c];
                for (int b = 0; b < a; ++b) v -= S[a * tB + b] * y[b];
                y[a] = v / S[a * tB + a];
            }
            for (int a = tB - 1; a >= 0; --a) {
                double v = y[a];
                for (int b = a + 1; b < tB; ++b) v -= S[b * tB + a] * X[b * D + c];
                X[a * D + c] = v / S[a * tB + a];
            }
        }
    }
    for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) {
            double acc = (r == c) ? 1.0 : 0.0;
            for (int a = 0; a < tB; ++a) acc -= L[a * D + r] * X[a * D + c];
            Sig[r * D + c] = acc;
        }
    out.Sig = Sig;
    std::vector<double> SB((size_t)D * std::max(u, 1)), CU((size_t)u * u);
    for (int r = 0; r < D; ++r)
        for (int a = 0; a < u; ++a) {
            double acc = 0.0;
            for (int c = 0; c < D; ++c) acc += Sig[r * D + c] * L[(tB + a) * D + c];
            SB[r * u + a] = acc;
        }
    for (int a = 0; a < u; ++a)
        for (int b = 0; b < u; ++b) {
            double acc = 0.0;
            for (int c = 0; c < D; ++c) acc += L[(tB + a) * D + c] * SB[c * u + b];
            CU[a * u + b] = acc;
        }
    out.CU = CU;
    out.BS.assign((size_t)u * D, 0.0);
    for (int a = 0; a < u; ++a)
        for (int r = 0; r < D; ++r) out.BS[a * D + r] = SB[r * u + a];
    out.mU.assign((size_t)G * std::max(u, 1), 0.0);
    out.mu.assign((size_t)G * D, 0.0);
    Nodes prior;                        // the tensor rule of ext serves every group of part 2
    if (D < kScFrom) prior = generate(D, m, L);
    // part 2 and the conditional moments, one group per relevance configuration of B (independent: over the host cores)
    parallel_for(G, 1, [&](int64_t g_lo, int64_t g_hi) {
        for (int64_t g = g_lo; g < g_hi; ++g) {
            // part 2: prior of ext inside (r_B = g, s*)
            {
                std::vector<double> e, w;
                const int b = (int)g | (out.sub_bits << tB);
                if (D < kScFrom) {
                    for (int32_t q = prior.group_begin[b]; q < prior.group_begin[b + 1]; ++q) {
                        for (int a = 0; a < D; ++a) e.push_back(prior.eta[(size_t)a * prior.n + q]);
                        w.push_back(prior.w[q]);
                    }
                } else {
                    orthant_nodes(D, m, L, b, e, w);
                }
                double acc = 0.0;
                for (double wk : w) acc += wk;
                out.mass[G + g] = acc;
                g_eta[G + g].swap(e);
                g_w[G + g].swap(w);
            }
            // moments conditional on the labels of B
            double* mu = out.mu.data() + (size_t)g * D;
            for (int c = 0; c < D; ++c) {
                double acc = 0.0;
                for (int a = 0; a < tB; ++a) acc += X[a * D + c] * ((((g >> a) & 1) ? 1.0 : -1.0) - m[a]);
                mu[c] = acc;
            }
            for (int a = 0; a < u; ++a) {
                double acc = m[tB + a];
                for (int c = 0; c < D; ++c) acc += L[(tB + a) * D + c] * mu[c];
                out.mU[(size_t)g * u + a] = acc;
            }
        }
    });
    // flatten, dimension-major
    out.group_begin.assign(2 * G + 1, 0);
    for (int g = 0; g < 2 * G; ++g) out.group_begin[g + 1] = out.group_begin[g] + (int32_t)g_w[g].size();
    out.n_nodes = out.group_begin[2 * G];
    out.eta.assign((size_t)D * out.n_nodes, 0.0);
    out.w.resize(out.n_nodes);
    for (int g = 0; g < 2 * G; ++g) {
        const int64_t pos = out.group_begin[g];
        for (size_t q = 0; q < g_w[g].size(); ++q) {
            out.w[pos + q] = g_w[g][q];
            for (int c = 0; c < D; ++c) out.eta[(size_t)c * out.n_nodes + pos + q] = g_eta[g][q * D + c];
        }
    }
    return out;
}

}  // namespace snq
