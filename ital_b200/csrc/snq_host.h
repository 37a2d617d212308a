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
#include <cmath>
#include <cstdint>
#include <vector>

namespace snq {

constexpr double kR = 7.0;
constexpr int kQMin = 2;
constexpr int kMaxOrder = 64;

inline int order_for(int t) {
    if (t <= 1) return 32;
    if (t == 2) return 16;
    if (t == 3) return 12;
    if (t == 4) return 6;
    return t == 5 ? 4 : 2;
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

// m[t], L[t*t] row-major lower triangular.
inline Nodes generate(int t, const double* m, const double* L, int q = 0, double R = kR) {
    Nodes out;
    out.t = t;
    if (q <= 0) q = order_for(t);
    const int two_q = 2 * q;
    int64_t n = 1;
    std::vector<double> eta(0), w(1, 1.0);
    std::vector<int32_t> orth(1, 0);
    const GaussLegendre& G = gl();
    for (int j = 0; j < t; ++j) {
        const int64_t n_new = n * two_q;
        std::vector<double> eta_new((size_t)(j + 1) * n_new), w_new(n_new);
        std::vector<int32_t> orth_new(n_new);
        for (int64_t k = 0; k < n; ++k) {
            double acc = m[j];
            for (int i = 0; i < j; ++i) acc += eta[(size_t)i * n + k] * L[j * t + i];
            const double a = -acc / L[j * t + j];
            const double c = a < -R ? -R : (a > R ? R : a);
            int n_lo = (int)std::floor(two_q * (c + R) / (2.0 * R) + 0.5);
            if (n_lo < kQMin) n_lo = kQMin;
            if (n_lo > two_q - kQMin) n_lo = two_q - kQMin;
            const int n_hi = two_q - n_lo;
            const double half_lo = 0.5 * (c + R), half_hi = 0.5 * (R - c);
            for (int g = 0; g < two_q; ++g) {
                const int64_t kk = k * two_q + g;
                double xg, wg;
                int bit;
                if (g < n_lo) {
                    xg = -R + half_lo * (1.0 + G.x[n_lo][g]);
                    wg = half_lo * G.w[n_lo][g] * phi(xg);
                    bit = 0;
                } else {
                    xg = c + half_hi * (1.0 + G.x[n_hi][g - n_lo]);
                    wg = half_hi * G.w[n_hi][g - n_lo] * phi(xg);
                    bit = 1;
                }
                for (int i = 0; i < j; ++i) eta_new[(size_t)i * n_new + kk] = eta[(size_t)i * n + k];
                eta_new[(size_t)j * n_new + kk] = xg;
                w_new[kk] = w[k] * wg;
                orth_new[kk] = orth[k] | (bit << j);
            }
        }
        eta.swap(eta_new);
        w.swap(w_new);
        orth.swap(orth_new);
        n = n_new;
    }
    // stable counting sort by orthant
    const int nb = 1 << t;
    out.n = n;
    out.group_begin.assign(nb + 1, 0);
    for (int64_t k = 0; k < n; ++k) out.group_begin[orth[k] + 1]++;
    for (int b = 0; b < nb; ++b) out.group_begin[b + 1] += out.group_begin[b];
    std::vector<int32_t> cursor(out.group_begin.begin(), out.group_begin.end() - 1);
    out.eta.resize((size_t)t * n);
    out.w.resize(n);
    out.orth.resize(n);
    for (int64_t k = 0; k < n; ++k) {
        const int32_t pos = cursor[orth[k]]++;
        for (int i = 0; i < t; ++i) out.eta[(size_t)i * n + pos] = eta[(size_t)i * n + k];
        out.w[pos] = w[k];
        out.orth[pos] = orth[k];
    }
    out.masses.assign(nb, 0.0);
    for (int b = 0; b < nb; ++b) {
        double s = 0.0;
        for (int32_t k = out.group_begin[b]; k < out.group_begin[b + 1]; ++k) s += out.w[k];
        out.masses[b] = s;
    }
    // same summand as the candidate scores (ital.py:207-219 with p' = 1), so that gain = score - entropy
    out.entropy = 0.0;
    const double eps = 1e-12, log1p_eps = std::log(1.0 + eps);
    for (int b = 0; b < nb; ++b) out.entropy += out.masses[b] * (log1p_eps - std::log(out.masses[b] + eps));
    return out;
}

}  // namespace snq
