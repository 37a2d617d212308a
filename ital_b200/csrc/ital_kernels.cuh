// Device kernels of the ITAL batch-selection path for sm_100a (B200).
//
// Memory layout in HBM (per shard, n = local rows, ldu = n rounded up to 32):
//   X      [n][d_pad]  float or double, row-major, rows zero-padded to a multiple of 512 bytes
//   sqn    [n]         double  |x_i|^2                                  (ital/gp.py:411,414-415)
//   m, v   [n]         double  posterior mean / variance given the labelled set (ital/gp.py:221-229)
//   U      [w_cap][ldu] double, COLUMN-major per projection: U[j][i] is entry j of L_K^-1 k(X_L, x_i),
//                      continued past the labelled set by the points already selected in the running batch
//                      (those extra columns are the l_i of SURVEY.md A.3)
//   gain   [n]         double  last exactly evaluated MI gain of row i in this fetch (lazy-greedy bound)
//   score  [n]         double  score of the last propose (NaN if not scored)
//   mask   [n]         uint8   1 seen, 2 selected in this fetch, 4 not a candidate
//
// Kernels:
//   k_sqnorm     one pass over X, |x_i|^2 in float64
//   k_extend     THE streaming pass: X row . z in float64, RBF value, projection against the new Cholesky
//                column, optional posterior mean/variance update.  HBM-bound: reads d_pad*sizeof(x) + 8*W + 8
//                bytes and writes 8 bytes per row (plus 32 bytes of m, v read-modify-write when labelling).
//   k_score0     closed-form first greedy step (one variable): binary entropy of Phi(m / sqrt(v))
//   k_eval       exact MI of the worklist candidates with the step's shared quadrature nodes, one warp each
//   k_argmax*    (score desc, index asc) reductions
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// tuning knobs of the streaming pass (see profiles/: chosen by measurement on B200)
#ifndef ITAL_EXTEND_RB
#define ITAL_EXTEND_RB 4
#endif
#ifndef ITAL_EXTEND_MINB
#define ITAL_EXTEND_MINB 2
#endif

namespace italk {

constexpr int kWarp = 32;
constexpr double kEps = 1e-12;           // MutualInformation eps (ital/ital.py:144)

struct Best {                            // result of an argmax reduction
    double score;
    long long idx;                       // local row, -1 if none
};

constexpr int kBaseStride = 16;          // row stride of the batch's Cholesky factor in device memory

struct CommitTargets {                   // where a single-shard step commits its winner (see k_pick_winner)
    double* rec_in = nullptr;
    double* rec_hist_t = nullptr;
    double* base_m = nullptr;
    double* base_L = nullptr;
    double* sel = nullptr;
    uint8_t* mask = nullptr;
    long long n = 0;
    int t = 0;
    int mark_bits = 0;
    int enabled = 0;
};

// Programmatic dependent launch: every kernel lets its successor in the stream start launching at once and then
// waits until its predecessor has completed and flushed (griddepcontrol.wait is a no-op for a launch without the
// programmatic-serialization attribute).  The stream order of all memory effects is unchanged; what is saved is the
// launch latency between the ~30 small dependent kernels of a fetch.
struct PickSrc {                         // what k_record reduces to find the step's local best (one of the two, or
    const Best* block_best = nullptr;    // neither: the row is given or `best` holds it): per-block results ...
    int nblocks = 0;
    const int* count = nullptr;          // ... or the scored rows of the final worklist
    const int* list = nullptr;
    const double* score = nullptr;
};

__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ bool better(double sa, long long ia, double sb, long long ib) {
    // NaN never wins; ties go to the lower index (np.argmax on the ascending candidate list, ital.py:98,130)
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    if (sa != sa) return false;
    if (sb != sb) return true;
    return sa > sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ double phi_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

// Per-row tag of the running fetch: [fetch epoch : 24 | batch columns valid : 4 | greedy step of the last exact score : 4].
// A tag written by an earlier fetch reads as "no batch column, not scored", so nothing has to be cleared between
// fetches (two n-sized memsets per fetch_unlabelled otherwise).  Steps are 1..15 (the first step is closed form and
// is never stamped), batch columns 0..15.
__device__ __forceinline__ int tag_ncol(uint32_t tag, uint32_t epoch) {
    return (tag >> 8) == epoch ? (int)((tag >> 4) & 15u) : 0;
}
__device__ __forceinline__ int tag_step(uint32_t tag, uint32_t epoch) {
    return (tag >> 8) == epoch ? (int)(tag & 15u) : 0;
}
__device__ __forceinline__ uint32_t tag_with_ncol(uint32_t tag, uint32_t epoch, int ncol) {
    return (epoch << 8) | ((uint32_t)ncol << 4) | (uint32_t)tag_step(tag, epoch);
}
__device__ __forceinline__ uint32_t tag_with_step(uint32_t tag, uint32_t epoch, int step) {
    return (epoch << 8) | ((uint32_t)tag_ncol(tag, epoch) << 4) | (uint32_t)step;
}

// barrier of a team of `n` threads (a multiple of 32) inside a block; id 0 with n = blockDim.x is __syncthreads()
__device__ __forceinline__ void team_barrier(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// Standard normal CDF from a table: (Phi, phi) on the grid x_k = -8.5 + k/128 and a 5th-order Taylor step around the
// nearest grid point (|delta| <= 1/256; derivatives of Phi are Hermite polynomials times phi).  Absolute error below
// 1e-16 against erfc, ~25 FP64 instructions and one 16-byte load instead of ~150 instructions; the quadrature sums
// need absolute, not relative accuracy.  The scoring kernels keep the table (35 KB) in shared memory: the lanes of a
// warp look up unrelated entries, which the L1 serves one cache line per cycle.
constexpr double kPhiXMax = 8.5;
constexpr int kPhiPerUnit = 128;
constexpr int kPhiTableLen = 17 * kPhiPerUnit + 1;             // 2177 grid points on [-8.5, 8.5]

__device__ __forceinline__ double phi_tab(const double2* tab, double x) {
    x = fmin(fmax(x, -kPhiXMax), kPhiXMax);                     // Phi(-8.5) = 1e-17, Phi(8.5) = 1 in double; NaN -> 0
    const int k = __double2int_rn((x + kPhiXMax) * kPhiPerUnit);
    const double xk = fma((double)k, 1.0 / kPhiPerUnit, -kPhiXMax);
    const double dl = x - xk;
    const double2 t = tab[k];
    const double x2 = xk * xk;
    const double c2 = -0.5 * xk;
    const double c3 = (x2 - 1.0) * (1.0 / 6.0);
    const double c4 = -xk * (x2 - 3.0) * (1.0 / 24.0);
    const double c5 = (x2 * (x2 - 6.0) + 3.0) * (1.0 / 120.0);
    const double poly = fma(dl, fma(dl, fma(dl, fma(dl, c5, c4), c3), c2), 1.0);
    return fma(t.y * dl, poly, t.x);
}

// copy of the table into shared memory (every thread of the block; followed by a block barrier at the caller)
__device__ __forceinline__ void phi_tab_to_shared(double2* dst, const double2* __restrict__ src) {
    for (int k = threadIdx.x; k < kPhiTableLen; k += blockDim.x) dst[k] = __ldg(src + k);
}

__device__ __forceinline__ double mi_term(double p, double log1p_eps) {
    // p * (log(p' + eps) - log(p + eps)) with p' = 1 (perfect user; ital.py:205-219)
    return p * (log1p_eps - log(p + kEps));
}

// Sum of 32 lane partials in the order of an xor butterfly (1, 2, 4, 8, 16): ((p0+p1)+(p2+p3))+...  Every kernel
// that sums lane partials of a row uses this association, whether the partials sit in shared memory (k_extend),
// travel through shuffles (k_catchup) or are transposed on the fly (k_extend_bulk), so the three agree bit for bit.
__device__ __forceinline__ double tree_sum32(const double* p) {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = p[2 * k] + p[2 * k + 1];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = a[2 * k] + a[2 * k + 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = a[2 * k] + a[2 * k + 1];
    return (a[0] + a[1]) + (a[2] + a[3]);
}

// One level of the on-the-fly transposition: `lo` belongs to the row group whose bit `o` is 0, `hi` to the one whose
// bit is 1.  Each lane hands the value of the other group to its partner (lane ^ o) and keeps the sum for its own.
__device__ __forceinline__ double tree_combine(double lo, double hi, int o, int lane) {
    const bool up = (lane & o) != 0;
    const double recv = __shfl_xor_sync(0xffffffffu, up ? lo : hi, o);
    return (up ? hi : lo) + recv;
}

template <typename XT> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    float4 v;
    __device__ __forceinline__ void load(const float* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    }
    __device__ __forceinline__ void lds(const float* p) { v = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ double get(int e) const {
        return (double)(e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w);
    }
};
template <> struct Vec<double> {
    static constexpr int N = 2;
    double2 v;
    __device__ __forceinline__ void load(const double* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    }
    __device__ __forceinline__ void lds(const double* p) { v = *reinterpret_cast<const double2*>(p); }
    __device__ __forceinline__ void zero() { v = make_double2(0.0, 0.0); }
    __device__ __forceinline__ double get(int e) const { return e == 0 ? v.x : v.y; }
};

// ---------------------------------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(256) k_sqnorm(const XT* __restrict__ X, int64_t n, int d_pad,
                                                double* __restrict__ sqn) {
    pdl_enter();
    constexpr int VN = Vec<XT>::N;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nchunks = d_pad / (32 * VN);
    for (int64_t row = warp; row < n; row += nwarps) {
        const XT* p = X + row * (int64_t)d_pad;
        double acc = 0.0;
        for (int c = 0; c < nchunks; ++c) {
            Vec<XT> x;
            x.load(p + (c * 32 + lane) * VN);
#pragma unroll
            for (int e = 0; e < VN; ++e) acc = fma(x.get(e), x.get(e), acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sqn[row] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// The streaming pass.  One warp owns a unit of 32 consecutive rows: phase 1 streams the rows with coalesced
// 16-byte loads (lane l holds columns (c*32 + l)*VN .. +VN of every row, the matching slice of z stays in
// registers) and leaves 32 partial sums per row in shared memory; phase 2 turns the warp by 90 degrees --
// lane l finishes row l -- so that the exp, the W-long projection and the stores run with all lanes busy and
// coalesced against the column-major U.
template <typename XT, int NC>   // NC: 16-byte chunks per lane and row (d_pad = NC * 32 * VN); 0 = run time
__global__ void __launch_bounds__(256, ITAL_EXTEND_MINB)
k_extend(const XT* __restrict__ X, int64_t n, int d, int d_pad, const double* __restrict__ rec, int w_cap, int W,
         const double* __restrict__ sqn, double* __restrict__ U, int64_t ldu,
         double* __restrict__ m, double* __restrict__ v, int labelled, double y, double noise, double var,
         double neg2ls2, uint8_t* __restrict__ mask, int64_t row_offset, uint8_t mark_bits) {
    pdl_enter();
    constexpr int VN = Vec<XT>::N;
    constexpr int RB = ITAL_EXTEND_RB;                  // rows in flight per warp
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int nwarp_blk = blockDim.x >> 5;
    double* part = smem + (size_t)wib * 32 * 33;        // [32 rows][33]
    double* ur_s = smem + (size_t)nwarp_blk * 32 * 33;  // [W]
    double* z_s = ur_s + ((W + 1) & ~1);                // [d_pad] only when NC == 0
    // the new point comes as a device-resident point record (header, projection, row as float64)
    if (rec[0] < 0.0) return;                           // empty record (no candidate was left): nothing to extend
    const double* ur = rec + 8;
    const double* z = rec + 8 + w_cap;
    if (mark_bits != 0 && blockIdx.x == 0 && threadIdx.x == 0) {    // the chosen row leaves the candidate set
        const long long loc = (long long)rec[0] - row_offset;
        if (loc >= 0 && loc < n) mask[loc] |= mark_bits;
    }
    for (int j = threadIdx.x; j < W; j += blockDim.x) ur_s[j] = ur[j];
    const int nchunks = NC > 0 ? NC : d_pad / (32 * VN);
    double zr[(NC > 0 ? NC : 1) * VN];
    if (NC > 0) {
#pragma unroll
        for (int c = 0; c < (NC > 0 ? NC : 1); ++c)
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const int col = (c * 32 + lane) * VN + e;
                zr[c * VN + e] = col < d ? z[col] : 0.0;
            }
    } else {
        for (int j = threadIdx.x; j < d_pad; j += blockDim.x) z_s[j] = j < d ? z[j] : 0.0;
    }
    __syncthreads();
    const double zn = rec[4];
    // labelled point: pivot of K_LL + noise I and the new entry of L_K^-1 y; batch point: noise-free pivot
    const double piv = labelled ? sqrt(fmax(rec[5] + noise, 2.3e-308)) : sqrt(fmax(rec[3], 1e-300));
    const double beta = labelled ? (y - rec[2]) / piv : 0.0;

    const int64_t n_units = (n + 31) >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * nwarp_blk + wib;
    const int64_t warps_total = (int64_t)gridDim.x * nwarp_blk;
    for (int64_t unit = warp_global; unit < n_units; unit += warps_total) {
        const int64_t row0 = unit << 5;
        // ---- phase 1: 32 partial dot products per row ----
        if (NC > 0) {
#pragma unroll 1
            for (int r = 0; r < 32; r += RB) {
                Vec<XT> x[RB][NC > 0 ? NC : 1];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const int64_t row = row0 + r + rr;
                    const XT* p = X + row * (int64_t)d_pad + lane * VN;
#pragma unroll
                    for (int c = 0; c < (NC > 0 ? NC : 1); ++c) {
                        if (row < n) x[rr][c].load(p + c * 32 * VN); else x[rr][c].zero();
                    }
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int c = 0; c < (NC > 0 ? NC : 1); ++c)
#pragma unroll
                        for (int e = 0; e < VN; e += 2) {
                            a0 = fma(x[rr][c].get(e), zr[c * VN + e], a0);
                            a1 = fma(x[rr][c].get(e + 1), zr[c * VN + e + 1], a1);
                        }
                    part[(r + rr) * 33 + lane] = a0 + a1;
                }
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < 32; r += RB) {
                double acc[RB];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) acc[rr] = 0.0;
                for (int c = 0; c < nchunks; ++c) {
                    Vec<XT> x[RB];
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) {
                        const int64_t row = row0 + r + rr;
                        if (row < n) x[rr].load(X + row * (int64_t)d_pad + (c * 32 + lane) * VN);
                        else x[rr].zero();
                    }
                    const double* zz = z_s + (c * 32 + lane) * VN;
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr)
#pragma unroll
                        for (int e = 0; e < VN; ++e) acc[rr] = fma(x[rr].get(e), zz[e], acc[rr]);
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) part[(r + rr) * 33 + lane] = acc[rr];
            }
        }
        __syncwarp();
        // ---- phase 2: lane l finishes row row0 + l ----
        const int64_t i = row0 + lane;
        if (i < n) {
            const double dot = tree_sum32(part + lane * 33);
            // v * exp((A + B - 2 * C) / s), s = -2 * sigma^2   (ital/gp.py:416)
            const double kv = var * exp((sqn[i] + zn - 2.0 * dot) / neg2ls2);
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            const double* u = U + i;
            int j = 0;
            for (; j + 4 <= W; j += 4) {
                p0 = fma(u[(int64_t)(j + 0) * ldu], ur_s[j + 0], p0);
                p1 = fma(u[(int64_t)(j + 1) * ldu], ur_s[j + 1], p1);
                p2 = fma(u[(int64_t)(j + 2) * ldu], ur_s[j + 2], p2);
                p3 = fma(u[(int64_t)(j + 3) * ldu], ur_s[j + 3], p3);
            }
            for (; j < W; ++j) p0 = fma(u[(int64_t)j * ldu], ur_s[j], p0);
            const double e = (kv - ((p0 + p1) + (p2 + p3))) / piv;
            U[(int64_t)W * ldu + i] = e;
            if (labelled) {
                m[i] = fma(e, beta, m[i]);
                v[i] = fma(-e, e, v[i]);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Streaming pass for Q labelled points at once (GaussianProcess.update with several samples, ital/gp.py:164-200):
// one read of X serves Q kernel columns -- the block-Cholesky extension of every row,
//   e_a = (k(x_i, z_a) - u_i . u_a - sum_{b<a} e_b T[a][b]) / piv_a,   a = 0..Q-1,
// with T the Q x Q triangle among the new points (computed on the host from their records), followed by
// m_i += sum_a e_a beta_a, v_i -= sum_a e_a^2.  The new rows sit in shared memory as float64; two rows of X are in
// flight per warp so that every shared-memory read of z is used twice.
struct MultiExt {                 // device block written by the host: header, then z[Q][d_pad], then ur[Q][W]
    double zn[4];
    double piv[4];
    double beta[4];
    double tri[16];               // tri[a * 4 + b], b < a
};

template <typename XT, int Q>
__global__ void __launch_bounds__(256, 2)
k_extend_multi(const XT* __restrict__ X, int64_t n, int d_pad, const double* __restrict__ ext, int W,
               const double* __restrict__ sqn, double* __restrict__ U, int64_t ldu, double* __restrict__ m,
               double* __restrict__ v, double var, double neg2ls2) {
    pdl_enter();
    constexpr int VN = Vec<XT>::N;
    extern __shared__ double msm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp_blk = blockDim.x >> 5;
    const MultiExt* hdr = reinterpret_cast<const MultiExt*>(ext);
    const double* z_g = ext + sizeof(MultiExt) / sizeof(double);
    const double* ur_g = z_g + (size_t)Q * d_pad;
    double* z_s = msm;                                  // [Q][d_pad]
    double* ur_s = z_s + (size_t)Q * d_pad;             // [Q][W]
    for (int j = threadIdx.x; j < Q * d_pad; j += blockDim.x) z_s[j] = z_g[j];
    for (int j = threadIdx.x; j < Q * W; j += blockDim.x) ur_s[j] = ur_g[j];
    __syncthreads();
    const int nchunks = d_pad / (32 * VN);
    const int64_t n_units = (n + 31) >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * nwarp_blk + wib;
    const int64_t warps_total = (int64_t)gridDim.x * nwarp_blk;
    for (int64_t unit = warp_global; unit < n_units; unit += warps_total) {
        const int64_t row0 = unit << 5;
        double dot[Q];
#pragma unroll
        for (int a = 0; a < Q; ++a) dot[a] = 0.0;
#pragma unroll 1
        for (int r = 0; r < 32; r += 2) {
            double acc0[Q], acc1[Q];
#pragma unroll
            for (int a = 0; a < Q; ++a) { acc0[a] = 0.0; acc1[a] = 0.0; }
            const int64_t ra = row0 + r, rb = row0 + r + 1;
            // chunks in groups of four: all eight 16-byte loads of a group are issued before any is used
            for (int c0 = 0; c0 < nchunks; c0 += 4) {
                Vec<XT> x0[4], x1[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = c0 + cc;
                    if (c < nchunks && ra < n) x0[cc].load(X + ra * (int64_t)d_pad + (c * 32 + lane) * VN); else x0[cc].zero();
                    if (c < nchunks && rb < n) x1[cc].load(X + rb * (int64_t)d_pad + (c * 32 + lane) * VN); else x1[cc].zero();
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = c0 + cc;
                    if (c < nchunks) {
                        double xd0[VN], xd1[VN];
#pragma unroll
                        for (int e = 0; e < VN; ++e) { xd0[e] = x0[cc].get(e); xd1[e] = x1[cc].get(e); }
#pragma unroll
                        for (int a = 0; a < Q; ++a) {
                            const double* zz = z_s + (size_t)a * d_pad + (c * 32 + lane) * VN;
#pragma unroll
                            for (int e = 0; e < VN; ++e) {
                                const double zv = zz[e];
                                acc0[a] = fma(xd0[e], zv, acc0[a]);
                                acc1[a] = fma(xd1[e], zv, acc1[a]);
                            }
                        }
                    }
                }
            }
            // lane sums by xor butterfly; lane r (r + 1) keeps the totals of its row
#pragma unroll
            for (int a = 0; a < Q; ++a) {
                double s0 = acc0[a], s1 = acc1[a];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                if (lane == r) dot[a] = s0;
                if (lane == r + 1) dot[a] = s1;
            }
        }
        const int64_t i = row0 + lane;
        if (i < n) {
            double proj[Q];
#pragma unroll
            for (int a = 0; a < Q; ++a) proj[a] = 0.0;
            const double* u = U + i;
            for (int j = 0; j < W; ++j) {
                const double uj = u[(int64_t)j * ldu];
#pragma unroll
                for (int a = 0; a < Q; ++a) proj[a] = fma(uj, ur_s[a * W + j], proj[a]);
            }
            const double sq = sqn[i];
            double e[Q];
            double dm = 0.0, dv = 0.0;
#pragma unroll
            for (int a = 0; a < Q; ++a) {
                double num = var * exp((sq + hdr->zn[a] - 2.0 * dot[a]) / neg2ls2) - proj[a];
#pragma unroll
                for (int b = 0; b < Q; ++b)
                    if (b < a) num = fma(-e[b], hdr->tri[a * 4 + b], num);
                e[a] = num / hdr->piv[a];
                U[(int64_t)(W + a) * ldu + i] = e[a];
                dm = fma(e[a], hdr->beta[a], dm);
                dv = fma(e[a], e[a], dv);
            }
            m[i] += dm;
            v[i] -= dv;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The streaming pass with the X stream staged through shared memory by the bulk-copy engine (TMA, 1-D
// cp.async.bulk + mbarrier): every warp owns a private ring of kBulkSlots slots of kBulkRows rows (4 KB at
// d = 512 float32); lane 0 issues one bulk copy per slot (the rows of a slot are contiguous in HBM) and the whole
// warp waits on the slot's mbarrier, so the loads need no registers and keep flowing while the warp is in its
// epilogue.  Arithmetic and its order are those of k_extend (bit-identical results).  NC > 0 only.
#ifndef ITAL_BULK_ROWS
#define ITAL_BULK_ROWS 1
#endif
#ifndef ITAL_BULK_SLOTS
#define ITAL_BULK_SLOTS 4
#endif
#ifndef ITAL_BULK_THREADS
#define ITAL_BULK_THREADS 512
#endif
constexpr int kBulkRows = ITAL_BULK_ROWS;
constexpr int kBulkSlots = ITAL_BULK_SLOTS;
constexpr int kBulkThreads = ITAL_BULK_THREADS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_issue(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <typename XT, int NC>
__global__ void __launch_bounds__(kBulkThreads, 1)
k_extend_bulk(const XT* __restrict__ X, int64_t n, int d, int d_pad, const double* __restrict__ rec, int w_cap, int W,
              const double* __restrict__ sqn, double* __restrict__ U, int64_t ldu,
              double* __restrict__ m, double* __restrict__ v, int labelled, double y, double noise, double var,
              double neg2ls2) {
    pdl_enter();
    constexpr int VN = Vec<XT>::N;
    extern __shared__ __align__(128) unsigned char bsm_raw[];
    if (rec[0] < 0.0) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp_blk = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)d_pad * sizeof(XT);
    const uint32_t slot_bytes = kBulkRows * row_bytes;
    // layout: [ring of every warp][ur][barriers]
    unsigned char* ring = bsm_raw + (size_t)wib * kBulkSlots * slot_bytes;
    double* ur_s = (double*)(bsm_raw + (size_t)nwarp_blk * kBulkSlots * slot_bytes);
    uint64_t* bars = (uint64_t*)(ur_s + ((W + 1) & ~1)) + wib * kBulkSlots;
    const double* ur = rec + 8;
    const double* z = rec + 8 + w_cap;
    for (int j = threadIdx.x; j < W; j += blockDim.x) ur_s[j] = ur[j];
    double zr[NC * VN];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int e = 0; e < VN; ++e) {
            const int col = (c * 32 + lane) * VN + e;
            zr[c * VN + e] = col < d ? z[col] : 0.0;
        }
    if (lane == 0) {
        for (int k = 0; k < kBulkSlots; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + k)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const double zn = rec[4];
    const double piv = labelled ? sqrt(fmax(rec[5] + noise, 2.3e-308)) : sqrt(fmax(rec[3], 1e-300));
    const double beta = labelled ? (y - rec[2]) / piv : 0.0;

    const int64_t n_units = (n + 31) >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * nwarp_blk + wib;
    const int64_t warps_total = (int64_t)gridDim.x * nwarp_blk;
    constexpr int kSubPerUnit = 32 / kBulkRows;
    // this warp's sequence of sub-tiles: (unit index in its own list) * kSubPerUnit + sub
    const int64_t my_units = warp_global < n_units ? (n_units - 1 - warp_global) / warps_total + 1 : 0;
    const int64_t total_sub = my_units * kSubPerUnit;
    auto issue = [&](int64_t sidx) {
        const int64_t unit = warp_global + (sidx / kSubPerUnit) * warps_total;
        const int64_t row = (unit << 5) + (sidx % kSubPerUnit) * kBulkRows;
        int64_t rows = n - row;
        if (rows > kBulkRows) rows = kBulkRows;
        const int slot = (int)(sidx % kBulkSlots);
        if (rows > 0)
            bulk_issue(ring + (size_t)slot * slot_bytes, X + row * (int64_t)d_pad, (uint32_t)rows * row_bytes, bars + slot);
        else    // nothing to load: complete the phase so that the waiters do not hang
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bars + slot)) : "memory");
    };
    if (lane == 0)
        for (int64_t k = 0; k < kBulkSlots && k < total_sub; ++k) issue(k);
    int64_t sidx = 0;
    for (int64_t unit = warp_global; unit < n_units; unit += warps_total) {
        const int64_t row0 = unit << 5;
        // Phase 1 with the 90-degree turn folded in: the lane partials of the 32 rows are transposed on the fly by a
        // binary tree of shuffles (pairs of rows, then pairs of pairs, ...), so that lane l ends up with the full
        // dot product of row l and no shared-memory tile is needed.  lv[k] holds the pending value of level k.
        double lv[5];
        double dot = 0.0;
#pragma unroll
        for (int sub = 0; sub < kSubPerUnit; ++sub, ++sidx) {
            const int slot = (int)(sidx % kBulkSlots);
            bar_wait(bars + slot, (uint32_t)((sidx / kBulkSlots) & 1));
            const unsigned char* sp = ring + (size_t)slot * slot_bytes;
#pragma unroll
            for (int rr = 0; rr < kBulkRows; ++rr) {
                const int r = sub * kBulkRows + rr;             // compile-time after unrolling
                const XT* xr = (const XT*)(sp + (size_t)rr * row_bytes) + lane * VN;
                double a0 = 0.0, a1 = 0.0;
                if (row0 + r < n) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        Vec<XT> x;
                        x.lds(xr + c * 32 * VN);
#pragma unroll
                        for (int e = 0; e < VN; e += 2) {
                            a0 = fma(x.get(e), zr[c * VN + e], a0);
                            a1 = fma(x.get(e + 1), zr[c * VN + e + 1], a1);
                        }
                    }
                }
                double val = a0 + a1;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (((r >> k) & 1) == 0) { lv[k] = val; break; }
                    val = tree_combine(lv[k], val, 1 << k, lane);
                    if (k == 4) dot = val;
                }
            }
            __syncwarp();                                   // every lane is done with the slot
            if (lane == 0 && sidx + kBulkSlots < total_sub) issue(sidx + kBulkSlots);
        }
        const int64_t i = row0 + lane;
        if (i < n) {
            const double kv = var * exp((sqn[i] + zn - 2.0 * dot) / neg2ls2);
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            const double* u = U + i;
            int j = 0;
            for (; j + 4 <= W; j += 4) {
                p0 = fma(u[(int64_t)(j + 0) * ldu], ur_s[j + 0], p0);
                p1 = fma(u[(int64_t)(j + 1) * ldu], ur_s[j + 1], p1);
                p2 = fma(u[(int64_t)(j + 2) * ldu], ur_s[j + 2], p2);
                p3 = fma(u[(int64_t)(j + 3) * ldu], ur_s[j + 3], p3);
            }
            for (; j < W; ++j) p0 = fma(u[(int64_t)j * ldu], ur_s[j], p0);
            const double e = (kv - ((p0 + p1) + (p2 + p3))) / piv;
            U[(int64_t)W * ldu + i] = e;
            if (labelled) {
                m[i] = fma(e, beta, m[i]);
                v[i] = fma(-e, e, v[i]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_extend_multi on the bulk-copy ring (2 KB rows): Q labelled points enter with one pass over the pool at close to
// the speed of the single-column pass.  Slots hold kMultiRows consecutive rows (one 8 KB bulk copy), so every
// shared-memory read of a new point's coordinates serves four rows (shared-memory bandwidth, not HBM, limits a
// two-row version); the new points sit in shared memory as float64, laid out so that the 32 lanes of a 16-byte
// read touch consecutive 16-byte words (no bank conflicts):
//   z_s[((a * NC + c) * (VN / 2) + h) * 64 + lane * 2 + e]  =  z_a[(c * 32 + lane) * VN + 2 h + e].
// The lane partials are transposed by the same shuffle tree as in k_extend_bulk.  8 warps per CTA, one CTA per SM;
// the ring keeps 8 warps x kMultiSlots x 8 KB in flight.
#ifndef ITAL_MULTI_THREADS
#define ITAL_MULTI_THREADS 256
#endif
#ifndef ITAL_MULTI_SLOTS
#define ITAL_MULTI_SLOTS 3
#endif
#ifndef ITAL_MULTI_ROWS
#define ITAL_MULTI_ROWS 4
#endif
constexpr int kMultiThreads = ITAL_MULTI_THREADS;
constexpr int kMultiSlots = ITAL_MULTI_SLOTS;
constexpr int kMultiRows = ITAL_MULTI_ROWS;

template <typename XT, int NC, int Q>
__global__ void __launch_bounds__(kMultiThreads, 1)
k_extend_bulk_multi(const XT* __restrict__ X, int64_t n, int d_pad, const double* __restrict__ ext, int W,
                    const double* __restrict__ sqn, double* __restrict__ U, int64_t ldu, double* __restrict__ m,
                    double* __restrict__ v, double var, double neg2ls2) {
    pdl_enter();
    constexpr int VN = Vec<XT>::N;
    constexpr int H = VN / 2;
    constexpr int R = kMultiRows;
    constexpr int kSubs = 32 / R;
    extern __shared__ __align__(128) unsigned char bmm_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp_blk = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)d_pad * sizeof(XT);
    const uint32_t slot_bytes = R * row_bytes;
    // layout: [ring of every warp][z_s][ur][barriers]
    unsigned char* ring = bmm_raw + (size_t)wib * kMultiSlots * slot_bytes;
    double* z_s = (double*)(bmm_raw + (size_t)nwarp_blk * kMultiSlots * slot_bytes);
    double* ur_s = z_s + (size_t)Q * d_pad;
    uint64_t* bars = (uint64_t*)(ur_s + ((Q * W + 1) & ~1)) + wib * kMultiSlots;
    const MultiExt* hdr = reinterpret_cast<const MultiExt*>(ext);
    const double* z_g = ext + sizeof(MultiExt) / sizeof(double);
    const double* ur_g = z_g + (size_t)Q * d_pad;
    for (int j = threadIdx.x; j < Q * d_pad; j += blockDim.x) {
        const int a = j / d_pad, col = j - a * d_pad;
        const int e = col % VN, ln = (col / VN) & 31, c = col / (VN * 32);
        z_s[((a * NC + c) * H + (e >> 1)) * 64 + ln * 2 + (e & 1)] = z_g[j];
    }
    for (int j = threadIdx.x; j < Q * W; j += blockDim.x) ur_s[j] = ur_g[j];
    if (lane == 0) {
        for (int k = 0; k < kMultiSlots; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + k)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int64_t n_units = (n + 31) >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * nwarp_blk + wib;
    const int64_t warps_total = (int64_t)gridDim.x * nwarp_blk;
    const int64_t my_units = warp_global < n_units ? (n_units - 1 - warp_global) / warps_total + 1 : 0;
    const int64_t total_sub = my_units * kSubs;
    auto issue = [&](int64_t sidx) {
        const int64_t unit = warp_global + (sidx / kSubs) * warps_total;
        const int64_t row = (unit << 5) + (sidx % kSubs) * R;
        int64_t rows = n - row;
        if (rows > R) rows = R;
        const int slot = (int)(sidx % kMultiSlots);
        if (rows > 0)
            bulk_issue(ring + (size_t)slot * slot_bytes, X + row * (int64_t)d_pad, (uint32_t)rows * row_bytes, bars + slot);
        else
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bars + slot)) : "memory");
    };
    if (lane == 0)
        for (int64_t k = 0; k < kMultiSlots && k < total_sub; ++k) issue(k);
    const double2* z2 = reinterpret_cast<const double2*>(z_s) + lane;
    int64_t sidx = 0;
    for (int64_t unit = warp_global; unit < n_units; unit += warps_total) {
        const int64_t row0 = unit << 5;
        double lv[Q][5];
        double dot[Q], proj[Q];
#pragma unroll
        for (int a = 0; a < Q; ++a) { dot[a] = 0.0; proj[a] = 0.0; }
        const int64_t i = row0 + lane;
        const double* u = U + i;
#pragma unroll
        for (int sub = 0; sub < kSubs; ++sub, ++sidx) {
            // the projection of this lane's row against the new points is folded into the main loop, four columns
            // of U per slot: the loads are issued here and used after the slot's arithmetic, so their latency is
            // covered (8 warps per SM cannot hide it in a separate epilogue)
            double up[4];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int col = sub * 4 + q4;
                up[q4] = (i < n && col < W) ? u[(int64_t)col * ldu] : 0.0;
            }
            const int slot = (int)(sidx % kMultiSlots);
            bar_wait(bars + slot, (uint32_t)((sidx / kMultiSlots) & 1));
            const XT* xr = (const XT*)(ring + (size_t)slot * slot_bytes) + lane * VN;
            double acc[R][Q];
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
#pragma unroll
                for (int a = 0; a < Q; ++a) acc[rr][a] = 0.0;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                double xd[R][VN];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                    Vec<XT> x;
                    x.lds(xr + (size_t)rr * d_pad + c * 32 * VN);
#pragma unroll
                    for (int e = 0; e < VN; ++e) xd[rr][e] = x.get(e);
                }
#pragma unroll
                for (int a = 0; a < Q; ++a) {
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        const double2 zz = z2[((a * NC + c) * H + h) * 32];
#pragma unroll
                        for (int rr = 0; rr < R; ++rr) acc[rr][a] = fma(xd[rr][2 * h], zz.x, acc[rr][a]);
#pragma unroll
                        for (int rr = 0; rr < R; ++rr) acc[rr][a] = fma(xd[rr][2 * h + 1], zz.y, acc[rr][a]);
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const int r = sub * R + rr;                     // compile-time after unrolling
                const bool valid = row0 + r < n;                // a slot may be partly filled at the end of the pool
#pragma unroll
                for (int a = 0; a < Q; ++a) {
                    double val = valid ? acc[rr][a] : 0.0;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        if (((r >> k) & 1) == 0) { lv[a][k] = val; break; }
                        val = tree_combine(lv[a][k], val, 1 << k, lane);
                        if (k == 4) dot[a] = val;
                    }
                }
            }
            __syncwarp();                                   // every lane is done with the slot
            if (lane == 0 && sidx + kMultiSlots < total_sub) issue(sidx + kMultiSlots);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int col = sub * 4 + q4;
                if (col < W) {
#pragma unroll
                    for (int a = 0; a < Q; ++a) proj[a] = fma(up[q4], ur_s[a * W + col], proj[a]);
                }
            }
        }
        if (i < n) {
            for (int j0 = 4 * kSubs; j0 < W; j0 += 8) {    // columns beyond the 32 covered above, eight loads in flight
                double ub[8];
#pragma unroll
                for (int q8 = 0; q8 < 8; ++q8) ub[q8] = j0 + q8 < W ? u[(int64_t)(j0 + q8) * ldu] : 0.0;
#pragma unroll
                for (int q8 = 0; q8 < 8; ++q8)
                    if (j0 + q8 < W) {
#pragma unroll
                        for (int a = 0; a < Q; ++a) proj[a] = fma(ub[q8], ur_s[a * W + j0 + q8], proj[a]);
                    }
            }
            const double sq = sqn[i];
            double e[Q];
            double dm = 0.0, dv = 0.0;
#pragma unroll
            for (int a = 0; a < Q; ++a) {
                double num = var * exp((sq + hdr->zn[a] - 2.0 * dot[a]) / neg2ls2) - proj[a];
#pragma unroll
                for (int b = 0; b < Q; ++b)
                    if (b < a) num = fma(-e[b], hdr->tri[a * 4 + b], num);
                e[a] = num / hdr->piv[a];
                U[(int64_t)(W + a) * ldu + i] = e[a];
                dm = fma(e[a], hdr->beta[a], dm);
                dv = fma(e[a], e[a], dv);
            }
            m[i] += dm;
            v[i] -= dv;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Lazy rows: bring the batch-conditional projection of the LISTED rows up to date (columns ncol .. t-1) from
// the stored records of the points selected so far, instead of streaming the whole pool once per greedy step
// (k_extend).  One warp per row; the arithmetic (per-lane partial sums, order of the lane sum, fma chains of the
// projection) is the same as in k_extend, so a row gets bit-identical entries either way.
//
// catchup_row: one warp, row i.  `recs` are the records of the selected points (stride rec_len; global or shared
// memory), `uv` is a warp-private scratch of w_cap doubles.  The row's tag keeps the number of valid batch columns.
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Everything the catch-up, the scoring and the record of row i will read, into L2 (fire and forget): the row of X, its
// projection entries, moments and tag.  Called by a whole warp for one row.
template <typename XT>
__device__ __forceinline__ void prefetch_row(int64_t i, int lane, const XT* X, int d_pad, const double* U, int64_t ldu,
                                             int n_cols, const double* m, const double* v, const double* sqn,
                                             const double* gain, const uint32_t* tags) {
    const char* xr = reinterpret_cast<const char*>(X + i * (int64_t)d_pad);
    const int lines = (int)((d_pad * sizeof(XT) + 127) / 128);
    for (int l = lane; l < lines; l += 32) prefetch_l2(xr + l * 128);
    for (int j = lane; j < n_cols; j += 32) prefetch_l2(U + (int64_t)j * ldu + i);
    if (lane == 0) prefetch_l2(m + i);
    if (lane == 1) prefetch_l2(v + i);
    if (lane == 2) prefetch_l2(sqn + i);
    if (lane == 3) prefetch_l2(gain + i);
    if (lane == 4) prefetch_l2(tags + i);
}

template <typename XT>
__device__ __forceinline__ void catchup_row(int64_t i, int lane, const XT* __restrict__ X, int d, int d_pad,
                                            const double* recs, int64_t rec_len, int w_cap, int W, int t,
                                            const double* __restrict__ sqn, double* U, int64_t ldu, uint32_t* tags,
                                            uint32_t epoch, double var, double neg2ls2, double* uv) {
    constexpr int VN = Vec<XT>::N;
    const int nchunks = d_pad / (32 * VN);
    const bool fixed = nchunks == 1 || nchunks == 2 || nchunks == 4;    // k_extend<XT, NC> vs k_extend<XT, 0>
    // every load the row needs is issued before the first one is used: the tag, the row of X (kept in registers for
    // the fixed shapes) and the projection entries up to the last column that can be valid (entries past the valid
    // ones are overwritten below before they are read)
    const uint32_t tag = __ldcg(tags + i);
    const XT* xrow = X + i * (int64_t)d_pad;
    Vec<XT> xr[4];
    if (fixed) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < nchunks) xr[c].load(xrow + (c * 32 + lane) * VN);
    }
    const double sq = sqn[i];
    for (int j = lane; j < W + t - 1; j += 32) uv[j] = __ldcg(U + (int64_t)j * ldu + i);
    const int c0 = tag_ncol(tag, epoch);
    if (c0 >= t) return;
    __syncwarp();
    for (int col = c0; col < t; ++col) {
        const double* rec = recs + (int64_t)col * rec_len;
        const double* z = rec + 8 + w_cap;
        double a0 = 0.0, a1 = 0.0;
        if (fixed) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < nchunks) {
#pragma unroll
                    for (int e = 0; e < VN; e += 2) {
                        const int cc = (c * 32 + lane) * VN + e;
                        const double z0 = cc < d ? z[cc] : 0.0, z1 = cc + 1 < d ? z[cc + 1] : 0.0;
                        a0 = fma(xr[c].get(e), z0, a0);
                        a1 = fma(xr[c].get(e + 1), z1, a1);
                    }
                }
            }
        } else {
            for (int c = 0; c < nchunks; ++c) {
                Vec<XT> x;
                x.load(xrow + (c * 32 + lane) * VN);
#pragma unroll
                for (int e = 0; e < VN; e += 2) {
                    const int cc = (c * 32 + lane) * VN + e;
                    const double z0 = cc < d ? z[cc] : 0.0, z1 = cc + 1 < d ? z[cc + 1] : 0.0;
                    a0 = fma(x.get(e), z0, a0);
                    a0 = fma(x.get(e + 1), z1, a0);
                }
            }
        }
        double dot = fixed ? a0 + a1 : a0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);   // = tree_sum32 order
        if (lane == 0) {
            const double kv = var * exp((sq + rec[4] - 2.0 * dot) / neg2ls2);
            const double* ur = rec + 8;
            const int Wc = W + col;
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            int j = 0;
            for (; j + 4 <= Wc; j += 4) {
                p0 = fma(uv[j + 0], ur[j + 0], p0);
                p1 = fma(uv[j + 1], ur[j + 1], p1);
                p2 = fma(uv[j + 2], ur[j + 2], p2);
                p3 = fma(uv[j + 3], ur[j + 3], p3);
            }
            for (; j < Wc; ++j) p0 = fma(uv[j], ur[j], p0);
            const double piv = sqrt(fmax(rec[3], 1e-300));
            const double e_new = (kv - ((p0 + p1) + (p2 + p3))) / piv;
            uv[Wc] = e_new;
            U[(int64_t)Wc * ldu + i] = e_new;
        }
        __syncwarp();
    }
    if (lane == 0) tags[i] = tag_with_ncol(tag, epoch, t);
    __syncwarp();
}

template <typename XT>
__global__ void __launch_bounds__(256) k_catchup(const int* __restrict__ count, const int* __restrict__ list,
                                                 const XT* __restrict__ X, int d, int d_pad,
                                                 const double* __restrict__ rec_hist, int64_t rec_len, int w_cap,
                                                 int W, int t, const double* __restrict__ sqn,
                                                 double* U, int64_t ldu, uint32_t* tags, uint32_t epoch,
                                                 double var, double neg2ls2) {
    pdl_enter();
    extern __shared__ double csm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp_blk = blockDim.x >> 5;
    double* uv = csm + (size_t)wib * w_cap;            // [w_cap] projection entries of the row
    const int n_items = *count;
    for (int item = blockIdx.x * nwarp_blk + wib; item < n_items; item += gridDim.x * nwarp_blk)
        catchup_row<XT>(list[item], lane, X, d, d_pad, rec_hist, rec_len, w_cap, W, t, sqn, U, ldu, tags, epoch, var,
                        neg2ls2, uv);
}

// Score of a single sample (first greedy step): MI of one variable in closed form (ital/ital.py:364-369, 183-224),
//   sum_r p_r (lc - log(p_r + eps)) = lc + H(u),  H(u) = -Phi(u) log(Phi(u) + eps) - Phi(-u) log(Phi(-u) + eps),
// u = |m| / sqrt(v) (H is even; p_0 + p_1 = 1).  H comes from a table of degree-7 polynomials on intervals of width
// 1/16 over [0, 8.5] (built on the host in extended precision, absolute error 1e-16 against erfc / log): ~40 FP64
// instructions per row instead of ~250, which is what a pass over 10^6 candidates is bound by.
constexpr int kHTabPerUnit = 16;
constexpr double kHTabMax = 8.5;
constexpr int kHTabLen = 136;                            // intervals; 8 coefficients each (ascending powers)

__device__ __forceinline__ double h_tab(const double* __restrict__ tab, double u) {
    u = fmin(u, kHTabMax);
    const int k = min((int)(u * kHTabPerUnit), kHTabLen - 1);
    const double x = fma(u, 2.0 * kHTabPerUnit, -(2.0 * k + 1.0));         // position in the interval, [-1, 1]
    const double2* c = reinterpret_cast<const double2*>(tab + 8 * k);
    const double2 c01 = __ldg(c), c23 = __ldg(c + 1), c45 = __ldg(c + 2), c67 = __ldg(c + 3);
    return fma(x, fma(x, fma(x, fma(x, fma(x, fma(x, fma(x, c67.y, c67.x), c45.y), c45.x), c23.y), c23.x), c01.y), c01.x);
}

__device__ __forceinline__ double score0_value(double mean, double var_raw, double lc, double scale,
                                               const double* __restrict__ htab) {
    const double var_i = fmax(var_raw, 0.0);             // predict_stored(cov_mode='diag') clamps (gp.py:229)
    const double sd = sqrt(var_i);
    const double u = sd > 0.0 ? fabs(mean) / sd : kHTabMax;
    if (u != u) return u;                                // a NaN mean stays NaN (and never wins)
    return scale * (lc + h_tab(htab, u));
}

// ---------------------------------------------------------------------------------------------------------
// First greedy step: one variable, closed form (ital/ital.py:364-369 and 183-224 with one configuration).
// Also seeds the lazy-greedy bound: gain = score.
__global__ void __launch_bounds__(256) k_score0(int64_t n, const double* __restrict__ m,
                                                const double* __restrict__ v, const uint8_t* __restrict__ mask,
                                                double* __restrict__ score, double* __restrict__ gain,
                                                Best* __restrict__ block_best, double log1p_eps, double scale,
                                                const double* __restrict__ htab) {
    pdl_enter();
    // perfect / mistaken user: scale = 1, log1p_eps = log(1 + eps) (a mistaken user adds a constant later);
    // general model (label_prob < 1): scale = label_prob, log1p_eps = (1-mp) log(1+eps) + mp log(eps)
    double bs = 0.0;
    long long bi = -1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 2 * stride) {
        // the loads of two rows are issued before the arithmetic of either
        uint8_t mk[2];
        double mm[2], vv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t i = i0 + u * stride;
            mk[u] = i < n ? __ldg(mask + i) : (uint8_t)1;
            mm[u] = i < n ? __ldg(m + i) : 0.0;
            vv[u] = i < n ? __ldg(v + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= n) break;
            double s = nan("");
            if (mk[u] == 0) {
                s = score0_value(mm[u], vv[u], log1p_eps, scale, htab);
                gain[i] = s;
                if (better(s, i, bs, bi)) { bs = s; bi = i; }
            }
            score[i] = s;
        }
    }
    // block reduction
    __shared__ double ss[32];
    __shared__ long long si[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(os, oi, bs, bi)) { bs = os; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = bs; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (better(ss[w], si[w], bs, bi)) { bs = ss[w]; bi = si[w]; }
        block_best[blockIdx.x].score = bs;
        block_best[blockIdx.x].idx = bi;
    }
}

// VarianceSampling (ital/baseline_methods.py:110-155): score of appending row i to the running batch,
//   sum of the variances minus sum of the covariances of ret + [i]  =  const(ret) + v_i - sum_a c(r_a, i),
// with c(r_a, i) = L_b[a] . l_i from the incremental Cholesky rows, i.e. v_i - g . l_i with g the column sums of L_b.
// t = 0 (and use_correlations = False): the posterior variance itself, clamped at 0 as predict_stored('diag') does.
// `allow_mask`: mask value that is admitted besides 0 (the reference's first pick does not exclude unnameable rows).
__global__ void __launch_bounds__(256) k_var_score(int64_t n, const double* __restrict__ v, const double* __restrict__ U,
                                                   int64_t ldu, int W0, int t, const double* __restrict__ base_L,
                                                   const uint8_t* __restrict__ mask, uint8_t allow_mask,
                                                   double* __restrict__ score, Best* __restrict__ block_best) {
    pdl_enter();
    __shared__ double g[16];
    if (threadIdx.x < 16) {
        double acc = 0.0;
        for (int a = threadIdx.x; a < t; ++a) acc += base_L[a * kBaseStride + threadIdx.x];
        g[threadIdx.x] = (int)threadIdx.x < t ? acc : 0.0;
    }
    __syncthreads();
    double bs = 0.0;
    long long bi = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t mk = mask[i];
        double s = nan("");
        if (mk == 0 || (allow_mask != 0 && mk == allow_mask)) {
            s = t == 0 ? fmax(v[i], 0.0) : v[i];
            for (int j = 0; j < t; ++j) s = fma(-g[j], U[(int64_t)(W0 + j) * ldu + i], s);
            if (better(s, i, bs, bi)) { bs = s; bi = i; }
        }
        score[i] = s;
    }
    __shared__ double ss[8];
    __shared__ long long si[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(os, oi, bs, bi)) { bs = os; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = bs; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (better(ss[w], si[w], bs, bi)) { bs = ss[w]; bi = si[w]; }
        block_best[blockIdx.x].score = bs;
        block_best[blockIdx.x].idx = bi;
    }
}

// argmax of `values` over candidate rows (mask == 0), per block
__global__ void __launch_bounds__(1024) k_argmax_rows(int64_t n, const double* __restrict__ values,
                                                      const uint8_t* __restrict__ mask,
                                                      Best* __restrict__ block_best, int* __restrict__ done,
                                                      int* __restrict__ count, int* __restrict__ list) {
    pdl_enter();
    double bs = 0.0;
    long long bi = -1;
    // four independent (mask, value) loads in flight per thread: the pass is latency-bound, not bandwidth-bound
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        uint8_t mk[4];
        double val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            mk[u] = i < n ? __ldg(mask + i) : (uint8_t)1;
            val[u] = i < n ? __ldg(values + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (mk[u] == 0 && better(val[u], i, bs, bi)) { bs = val[u]; bi = i; }
        }
    }
    __shared__ double ss[32];
    __shared__ long long si[32];
    __shared__ int last_s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(os, oi, bs, bi)) { bs = os; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = bs; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (better(ss[w], si[w], bs, bi)) { bs = ss[w]; bi = si[w]; }
        block_best[blockIdx.x].score = bs;
        block_best[blockIdx.x].idx = bi;
        // the last block to finish turns the per-block winners into the stage-A worklist (fixed order)
        __threadfence();
        last_s = (list != nullptr && atomicAdd(done, 1) == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (last_s) {                                       // (*count is 0 on entry; the order of the list is immaterial)
        __threadfence();
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
            const long long idx = block_best[k].idx;
            if (idx >= 0) list[atomicAdd(count, 1)] = (int)idx;
        }
        if (threadIdx.x == 0) *done = 0;
    }
}

// argmax of score[] over the rows listed in the worklist, per block (k_record reduces the per-block results)
__global__ void __launch_bounds__(256) k_argmax_list(const int* __restrict__ count,
                                                     const int* __restrict__ list,
                                                     const double* __restrict__ score,
                                                     Best* __restrict__ block_best,     // gridDim.x == 1: final
                                                     const double* __restrict__ h_base = nullptr,
                                                     double floor_score = 0.0, double margin = 0.0,
                                                     double* __restrict__ thr_gain = nullptr,
                                                     int* __restrict__ reset_count = nullptr) {
    pdl_enter();
    const int n = *count;
    double bs = 0.0;
    long long bi = -1;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const long long i = list[k];
        if (better(score[i], i, bs, bi)) { bs = score[i]; bi = i; }
    }
    __shared__ double ss[32];
    __shared__ long long si[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(os, oi, bs, bi)) { bs = os; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = bs; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (better(ss[w], si[w], bs, bi)) { bs = ss[w]; bi = si[w]; }
        block_best[blockIdx.x].score = bs;
        block_best[blockIdx.x].idx = bi;
        if (thr_gain != nullptr) {      // stage B threshold: a row can still win only if its bound reaches this
            double thr = floor_score;
            if (bi >= 0 && bs == bs) thr = fmax(thr, bs);
            *thr_gain = thr - margin - *h_base;
            *reset_count = 0;           // the worklist is rebuilt next
        }
    }
}

// Lazy-greedy worklist: candidate rows whose upper bound gain_i reaches *thr_gain (a device scalar written by
// k_argmax_list at the end of stage A); every candidate when exhaustive.  Rows already scored in this
// step are left out unless `keep_scored` (the final list must contain them for the argmax).
__global__ void __launch_bounds__(256) k_worklist(int64_t n, const uint8_t* __restrict__ mask,
                                                  const double* __restrict__ gain,
                                                  const double* __restrict__ thr_gain, int exhaustive,
                                                  int* __restrict__ count, int* __restrict__ list) {
    pdl_enter();
    const double thr = exhaustive ? -1e300 : *thr_gain;
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n;
         i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane;
        const uint8_t mk = i < n ? __ldg(mask + i) : (uint8_t)1;      // both loads issued before either is used
        const double gv = i < n ? __ldg(gain + i) : 0.0;
        // (exhaustive: the gain is not looked at -- it may never have been written, and a NaN would drop the row)
        const bool take = mk == 0 && (exhaustive != 0 || gv >= thr);
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        if (ballot != 0) {
            int base = 0;
            if (lane == 0) base = atomicAdd(count, __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (take) list[base + __popc(ballot & ((1u << lane) - 1))] = (int)i;
        }
    }
}

// Worklist of ONE given row (already stored in list[0]) if it is a candidate.
__global__ void k_worklist_one(const uint8_t* __restrict__ mask, const int* __restrict__ list, int* __restrict__ count) {
    pdl_enter();
    if (threadIdx.x == 0) count[0] = mask[list[0]] == 0 ? 1 : 0;
}

// Exact MI of one candidate per team of threads (a warp, or the whole 256-thread block when few candidates are
// left and latency matters) with the shared nodes of the step (t >= 1 base variables):
//   P(r_base, +) = sum_{q in orthant r_base} w_q * Phi((m_i + l_i . eta_q) / s_i),  P(r_base, -) = P(r_base) - P(+)
//   score = sum_r p_r * (log(1 + eps) - log(p_r + eps))
// Replaces 2^(t+1) calls of prob_rel + updated_prob_rel per candidate (ital/ital.py:193-219).  The reduction
// order is fixed, so identical rows get bit-identical scores.  Rows already scored in this step are skipped.
//
// k_eval<T>, T = 1..3: nodes in generation order with their orthant id (k_snq_generate), 2^T accumulators per
// thread.  k_eval_sorted: any t, nodes sorted by orthant on the host (t >= 4).
struct EvalArgs {
    const int* count;
    const int* list;
    const double* m;
    const double* v;
    const double* U;
    int64_t ldu;
    int W0;
    const double* eta;          // k_eval_sorted: [t][n_nodes]
    const double* w;            // k_eval_sorted
    const double* nodes4;       // k_eval<T>: {eta_0, eta_1, eta_2, weight} per node, sorted by orthant, every orthant
                                // zero-padded to a multiple of kNodePad nodes; group_begin[2^T + 1] = offsets
    const double2* phi;         // table of phi_tab
    const int* orth;            // (orthant id per generated node: k_snq_generate -> k_snq_finalize)
    const int* group_begin;     // 2^t + 1 offsets of the orthants in the node list
    int64_t n_nodes;            // stride of eta (nodes generated)
    const int* n_kept;          // nodes kept after dropping negligible weights (device; k_eval<T>)
    const double* masses;       // 2^t base orthant probabilities (device)
    const double* h_base;       // score of the base alone (device scalar)
    double log1p_eps;
    double flag_var;
    double* score;
    double* gain;
    uint32_t* tags;             // tag_step(tags[i]) == t: row i has been scored in this greedy step
    uint32_t epoch;
    int* n_flagged;
    int* n_scored;
    int force_block;
    int t;
};

__device__ __forceinline__ int eval_team_size(int n_items, int force_block) {
    // a warp per candidate when there are enough candidates to keep every warp busy, otherwise the whole block
    // works on one candidate (the node loop is latency-bound for a lone warp)
    return (force_block || n_items < (int)(gridDim.x * (blockDim.x >> 5))) ? (int)blockDim.x : 32;
}

constexpr int kNodePad = 256;            // every orthant's nodes are zero-padded to a multiple of this (one node per thread of a team)

__host__ __device__ __forceinline__ int pad_nodes(int cnt) { return (cnt + kNodePad - 1) / kNodePad * kNodePad; }

// Exact score of candidate i by a team of TPC threads (32, or 256 = eight warps synchronising on barrier `bar_id`)
// with the nodes of the step: {eta, weight} packed per node, sorted by orthant in generation order, every orthant
// padded with zero-weight nodes to a multiple of 256 (no tail, no orthant id per node, one accumulator per thread).
// Thread `tid_team` takes the nodes g0 + tid_team + TPC j of an orthant in that order; the partial sums are reduced
// by an xor butterfly inside every warp and then warp by warp in ascending order, so a row's score depends on the
// team size only.  `red` holds 8 * 2^T doubles per team; `phi` is the table of phi_tab (shared memory); `gb` the
// orthant offsets and `masses` the base orthant masses (shared or global memory).  Returns after the row's score,
// gain and tag are written.
// The tail of a candidate's score: the per-warp orthant sums `acc` are added over the team in a fixed order, lane b
// finishes orthant b, lane 0 adds the terms in ascending order of b and writes score, gain and tag.
template <int T>
__device__ __forceinline__ void eval_epilogue(const EvalArgs& a, int64_t i, int tid_team, int TPC, int bar_id, double* red,
                                              const double* masses, double h_base, const double* acc, double s2) {
    constexpr int NB = 1 << T;
    if (TPC > 32) {
        if ((tid_team & 31) == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) red[b * 8 + (tid_team >> 5)] = acc[b];
        }
        team_barrier(bar_id, TPC);
    }
    if (tid_team < 32) {
        // lane b of the team's first warp finishes orthant b (its two logarithms); lane 0 adds the terms in
        // ascending order of b
        double p_plus = 0.0;
        if (TPC > 32) {
            if (tid_team < NB)
                for (int k = 0; k < TPC / 32; ++k) p_plus += red[tid_team * 8 + k];
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) p_plus = (tid_team == b) ? acc[b] : p_plus;
        }
        double term = 0.0;
        if (tid_team < NB) {
            const double p_minus = fmax(masses[tid_team] - p_plus, 0.0);
            term = mi_term(p_plus, a.log1p_eps) + mi_term(p_minus, a.log1p_eps);
        }
        double sc = 0.0;
#pragma unroll
        for (int b = 0; b < NB; ++b) sc += __shfl_sync(0xffffffffu, term, b);
        if (tid_team == 0) {
            a.tags[i] = tag_with_step(__ldcg(a.tags + i), a.epoch, a.t);
            a.score[i] = sc;
            a.gain[i] = sc - h_base;
            atomicAdd(a.n_scored, 1);
            if (s2 < a.flag_var) atomicAdd(a.n_flagged, 1);
        }
    }
    if (TPC > 32) team_barrier(bar_id, TPC);            // red and the tag are free again
}

template <int T>
__device__ __forceinline__ void eval_candidate(const EvalArgs& a, int64_t i, int tid_team, int TPC, int bar_id,
                                               double* red, const double2* phi, const int* gb, const double* masses,
                                               double h_base) {
    constexpr int NB = 1 << T;
    double l[3] = {0.0, 0.0, 0.0};
    double s2 = a.v[i];
#pragma unroll
    for (int j = 0; j < T; ++j) {
        l[j] = __ldcg(a.U + (int64_t)(a.W0 + j) * a.ldu + i);
        s2 = fma(-l[j], l[j], s2);
    }
    const double mi = a.m[i];
    const double s = s2 > 0.0 ? sqrt(s2) : 0.0;
    const double inv_s = s > 0.0 ? 1.0 / s : 0.0;
    const double2* nd = reinterpret_cast<const double2*>(a.nodes4);
    double acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int g0 = gb[b], g1 = gb[b + 1];
        double ac = 0.0;
        if (s > 0.0) {
#pragma unroll 4
            for (int q = g0 + tid_team; q < g1; q += TPC) {
                const double2 n01 = nd[2 * q], n23 = nd[2 * q + 1];
                double num = fma(l[0], n01.x, mi);
                if (T >= 2) num = fma(l[1], n01.y, num);
                if (T >= 3) num = fma(l[2], n23.x, num);
                ac = fma(n23.y, phi_tab(phi, num * inv_s), ac);
            }
        } else {                                        // no conditional variance left: Phi is a step
            for (int q = g0 + tid_team; q < g1; q += TPC) {
                const double2 n01 = nd[2 * q], n23 = nd[2 * q + 1];
                double num = fma(l[0], n01.x, mi);
                if (T >= 2) num = fma(l[1], n01.y, num);
                if (T >= 3) num = fma(l[2], n23.x, num);
                ac = fma(n23.y, num > 0.0 ? 1.0 : 0.0, ac);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ac += __shfl_xor_sync(0xffffffffu, ac, o);
        acc[b] = ac;
    }
    eval_epilogue<T>(a, i, tid_team, TPC, bar_id, red, masses, h_base, acc, s2);
}

// Two candidates per warp against ONE pass over the node list (exhaustive scoring, where there are far more candidates
// than warps): every node is loaded once and feeds two independent Phi evaluations, which halves the L2 traffic of the
// node list (262 KB per pass at three base variables, more than the L1 keeps) and doubles the work in flight per
// load.  Each candidate's sums are formed in the same order as in eval_candidate: bit-identical scores.
template <int T>
__device__ __forceinline__ void eval_candidate_pair(const EvalArgs& a, int64_t i0, int64_t i1, int lane, const double2* phi,
                                                    const int* gb, const double* masses, double h_base) {
    constexpr int NB = 1 << T;
    double l0[3] = {0.0, 0.0, 0.0}, l1[3] = {0.0, 0.0, 0.0};
    double s20 = a.v[i0], s21 = a.v[i1];
#pragma unroll
    for (int j = 0; j < T; ++j) {
        l0[j] = __ldcg(a.U + (int64_t)(a.W0 + j) * a.ldu + i0);
        l1[j] = __ldcg(a.U + (int64_t)(a.W0 + j) * a.ldu + i1);
        s20 = fma(-l0[j], l0[j], s20);
        s21 = fma(-l1[j], l1[j], s21);
    }
    const double m0 = a.m[i0], m1 = a.m[i1];
    const double s0 = s20 > 0.0 ? sqrt(s20) : 0.0, s1 = s21 > 0.0 ? sqrt(s21) : 0.0;
    const double inv0 = s0 > 0.0 ? 1.0 / s0 : 0.0, inv1 = s1 > 0.0 ? 1.0 / s1 : 0.0;
    const double2* nd = reinterpret_cast<const double2*>(a.nodes4);
    double acc0[NB], acc1[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int g0 = gb[b], g1 = gb[b + 1];
        double ac0 = 0.0, ac1 = 0.0;
#pragma unroll 2
        for (int q = g0 + lane; q < g1; q += 32) {
            const double2 n01 = nd[2 * q], n23 = nd[2 * q + 1];
            double num0 = fma(l0[0], n01.x, m0), num1 = fma(l1[0], n01.x, m1);
            if (T >= 2) { num0 = fma(l0[1], n01.y, num0); num1 = fma(l1[1], n01.y, num1); }
            if (T >= 3) { num0 = fma(l0[2], n23.x, num0); num1 = fma(l1[2], n23.x, num1); }
            const double c0 = s0 > 0.0 ? phi_tab(phi, num0 * inv0) : (num0 > 0.0 ? 1.0 : 0.0);
            const double c1 = s1 > 0.0 ? phi_tab(phi, num1 * inv1) : (num1 > 0.0 ? 1.0 : 0.0);
            ac0 = fma(n23.y, c0, ac0);
            ac1 = fma(n23.y, c1, ac1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ac0 += __shfl_xor_sync(0xffffffffu, ac0, o);
            ac1 += __shfl_xor_sync(0xffffffffu, ac1, o);
        }
        acc0[b] = ac0;
        acc1[b] = ac1;
    }
    eval_epilogue<T>(a, i0, lane, 32, 0, nullptr, masses, h_base, acc0, s20);
    eval_epilogue<T>(a, i1, lane, 32, 0, nullptr, masses, h_base, acc1, s21);
}

template <int T>
__global__ void __launch_bounds__(256) k_eval(EvalArgs a) {
    pdl_enter();
    constexpr int NB = 1 << T;
    const int n_items = *a.count;
    const int TPC = eval_team_size(n_items, a.force_block);
    const int tid_team = threadIdx.x % TPC;
    const int team_global = (blockIdx.x * blockDim.x + threadIdx.x) / TPC;
    const int teams_total = (gridDim.x * blockDim.x) / TPC;
    const double h_base = *a.h_base;
    __shared__ double red[8 * NB];
    __shared__ int gb[NB + 1];
    __shared__ double2 phi_s[kPhiTableLen];
    if (blockIdx.x * (blockDim.x / TPC) >= n_items) return;     // no item for this block
    phi_tab_to_shared(phi_s, a.phi);
    if (threadIdx.x <= NB) gb[threadIdx.x] = a.group_begin[threadIdx.x];
    __syncthreads();
    if (TPC == 32 && n_items >= 2 * teams_total) {
        // far more candidates than warps: two per warp and pass over the node list
        for (int p = team_global; 2 * p < n_items; p += teams_total) {
            const int64_t i0 = a.list[2 * p];
            const int64_t i1 = 2 * p + 1 < n_items ? a.list[2 * p + 1] : -1;
            const bool do0 = tag_step(a.tags[i0], a.epoch) != a.t;
            const bool do1 = i1 >= 0 && tag_step(a.tags[i1], a.epoch) != a.t;
            if (do0 && do1) eval_candidate_pair<T>(a, i0, i1, tid_team, phi_s, gb, a.masses, h_base);
            else if (do0) eval_candidate<T>(a, i0, tid_team, TPC, 0, red, phi_s, gb, a.masses, h_base);
            else if (do1) eval_candidate<T>(a, i1, tid_team, TPC, 0, red, phi_s, gb, a.masses, h_base);
        }
        return;
    }
    for (int item = team_global; item < n_items; item += teams_total) {
        const int64_t i = a.list[item];
        if (tag_step(a.tags[i], a.epoch) == a.t) continue;  // scored earlier in this step (team-uniform)
        eval_candidate<T>(a, i, tid_team, TPC, 0, red, phi_s, gb, a.masses, h_base);
    }
}

__device__ __forceinline__ double team_sum(double acc, int TPC, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (TPC > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        acc = 0.0;
        for (int k = 0; k < TPC / 32; ++k) acc += red[k];
    }
    return acc;
}

// More candidates than teams (the wide pruning margins of the long-batch rules let thousands of rows through): two
// candidates per team and pass over the node list.  The list (up to 30 MB) streams from L2 once per pass, which is what
// bounds the scoring of long batches; each candidate's sums are formed in the same order as in k_eval_sorted.  A
// kernel of its own so that the extra registers do not cost k_eval_sorted its occupancy.
__global__ void __launch_bounds__(256) k_eval_sorted_pair(EvalArgs a) {
    pdl_enter();
    constexpr int MAXT = 10;
    const int t = a.t;
    const int n_items = *a.count;
    const int TPC = eval_team_size(n_items, a.force_block);
    const int tid_team = threadIdx.x % TPC;
    const int team_global = (blockIdx.x * blockDim.x + threadIdx.x) / TPC;
    const int teams_total = (gridDim.x * blockDim.x) / TPC;
    const int64_t N = a.n_nodes;
    __shared__ double red[8];
    __shared__ double2 phi_s[kPhiTableLen];
    phi_tab_to_shared(phi_s, a.phi);
    __syncthreads();
    if (n_items <= teams_total) return;                // k_eval_sorted takes these
    {
        for (int p = team_global; 2 * p < n_items; p += teams_total) {
            int64_t ii[2] = {a.list[2 * p], 2 * p + 1 < n_items ? a.list[2 * p + 1] : -1};
            bool todo[2];
            double l[2][MAXT], s2[2], mi[2], inv_s[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                todo[c] = ii[c] >= 0 && tag_step(a.tags[ii[c]], a.epoch) != a.t;
                const int64_t i = todo[c] ? ii[c] : a.list[2 * p];
                s2[c] = a.v[i];
#pragma unroll
                for (int j = 0; j < MAXT; ++j) {
                    l[c][j] = 0.0;
                    if (j < t) {
                        l[c][j] = a.U[(int64_t)(a.W0 + j) * a.ldu + i];
                        s2[c] = fma(-l[c][j], l[c][j], s2[c]);
                    }
                }
                mi[c] = a.m[i];
                const double sd = s2[c] > 0.0 ? sqrt(s2[c]) : 0.0;
                inv_s[c] = sd > 0.0 ? 1.0 / sd : 0.0;
            }
            if (!todo[0] && !todo[1]) continue;
            double sc[2] = {0.0, 0.0};
            const int nb = 1 << t;
            for (int b = 0; b < nb; ++b) {
                const int g0 = a.group_begin[b], g1 = a.group_begin[b + 1];
                double acc[2] = {0.0, 0.0};
                for (int q = g0 + tid_team; q < g1; q += TPC) {
                    double num[2] = {mi[0], mi[1]};
#pragma unroll
                    for (int j = 0; j < MAXT; ++j)
                        if (j < t) {
                            const double ej = a.eta[(int64_t)j * N + q];
                            num[0] = fma(l[0][j], ej, num[0]);
                            num[1] = fma(l[1][j], ej, num[1]);
                        }
                    const double wq = a.w[q];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const double cdf = inv_s[c] > 0.0 ? phi_tab(phi_s, num[c] * inv_s[c]) : (num[c] > 0.0 ? 1.0 : 0.0);
                        acc[c] = fma(wq, cdf, acc[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double p_plus = team_sum(acc[c], TPC, red);
                    const double p_minus = fmax(a.masses[b] - p_plus, 0.0);
                    sc[c] += mi_term(p_plus, a.log1p_eps) + mi_term(p_minus, a.log1p_eps);
                }
            }
            if (TPC > 32) __syncthreads();
            if (tid_team == 0) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (todo[c]) {
                        const int64_t i = ii[c];
                        a.tags[i] = tag_with_step(a.tags[i], a.epoch, a.t);
                        a.score[i] = sc[c];
                        a.gain[i] = sc[c] - *a.h_base;
                        atomicAdd(a.n_scored, 1);
                        if (s2[c] < a.flag_var) atomicAdd(a.n_flagged, 1);
                    }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_eval_sorted(EvalArgs a) {
    pdl_enter();
    constexpr int MAXT = 10;
    const int t = a.t;
    const int n_items = *a.count;
    const int TPC = eval_team_size(n_items, a.force_block);
    const int tid_team = threadIdx.x % TPC;
    const int team_global = (blockIdx.x * blockDim.x + threadIdx.x) / TPC;
    const int teams_total = (gridDim.x * blockDim.x) / TPC;
    const int64_t N = a.n_nodes;
    __shared__ double red[8];
    __shared__ double2 phi_s[kPhiTableLen];
    phi_tab_to_shared(phi_s, a.phi);
    __syncthreads();
    if (n_items > teams_total) return;                 // k_eval_sorted_pair takes these
    for (int item = team_global; item < n_items; item += teams_total) {
        const int64_t i = a.list[item];
        if (tag_step(a.tags[i], a.epoch) == a.t) continue;
        double l[MAXT];
        double s2 = a.v[i];
#pragma unroll
        for (int j = 0; j < MAXT; ++j) {
            l[j] = 0.0;
            if (j < t) {
                l[j] = a.U[(int64_t)(a.W0 + j) * a.ldu + i];
                s2 = fma(-l[j], l[j], s2);
            }
        }
        const double mi = a.m[i];
        const double s = s2 > 0.0 ? sqrt(s2) : 0.0;
        const double inv_s = s > 0.0 ? 1.0 / s : 0.0;
        double sc = 0.0;
        const int nb = 1 << t;
        for (int b = 0; b < nb; ++b) {
            const int g0 = a.group_begin[b], g1 = a.group_begin[b + 1];
            double acc = 0.0;
            for (int q = g0 + tid_team; q < g1; q += TPC) {
                double num = mi;
#pragma unroll
                for (int j = 0; j < MAXT; ++j)
                    if (j < t) num = fma(l[j], a.eta[(int64_t)j * N + q], num);
                const double cdf = s > 0.0 ? phi_tab(phi_s, num * inv_s) : (num > 0.0 ? 1.0 : 0.0);
                acc = fma(a.w[q], cdf, acc);
            }
            const double p_plus = team_sum(acc, TPC, red);
            const double p_minus = fmax(a.masses[b] - p_plus, 0.0);
            sc += mi_term(p_plus, a.log1p_eps) + mi_term(p_minus, a.log1p_eps);
        }
        if (TPC > 32) __syncthreads();
        if (tid_team == 0) {
            a.tags[i] = tag_with_step(a.tags[i], a.epoch, a.t);
            a.score[i] = sc;
            a.gain[i] = sc - *a.h_base;
            atomicAdd(a.n_scored, 1);
            if (s2 < a.flag_var) atomicAdd(a.n_flagged, 1);
        }
    }
}

// General feedback model (label_prob < 1): MI of one candidate per block from the conditional node sets of
// csrc/snq_host.h (generate_general).  For every (set, orthant of the unlabelled base variables) it accumulates
//   A  = sum w Phi((m_i + l_i.eta)/s_i)                        candidate unlabelled, relevant
//   B+ = sum w exp(-((+1 - m_i - l_i.eta)/s~_i)^2 / 2)          candidate labelled relevant   (s~^2 = s^2 + noise)
//   B- = sum w exp(-((-1 - m_i - l_i.eta)/s~_i)^2 / 2)          candidate labelled irrelevant
// and then assembles  MI = sum_r p_r { sum_{O != 0} (1-lp)^(D-|O|) lp^|O| [ (1-mp)^|O| log(q_{r,O} + eps)
//                                      + (1 - (1-mp)^|O|) log eps ] - (1 - (1-lp)^D) log(p_r + eps) }
// (MutualInformation._call_iter_all with fb_iter's general case, ital/ital.py:183-224, 330-342, 453-481; q_{r,O} is
// updated_prob_rel, ital/ital.py:432-450, with the labelled samples pinned -- DESIGN.md "general feedback model").
struct GeneralArgs {
    const int* count;
    const int* list;
    const double* m;
    const double* v;
    const double* U;
    int64_t ldu;
    int W0;
    int t;
    const double* eta;
    const double* w;
    int64_t n_nodes;
    const int* group_begin;
    const double* group_mass;
    const int* set_group0;
    const int* lut;
    int n_groups;
    int n_sets;
    double lp, mp, noise;
    const double2* phi;
    double* score;
    double* gain;
    uint32_t* tags;
    uint32_t epoch;
    int* n_scored;
    int estimation;             // label_estimation: 0 'mean', 1 'optimistic', 2 'pessimistic' (ital.py:210-219)
    int fb_kind;                // feedback configurations enumerated (fb_iter, ital.py:300-342): 0 the single perfect
                                // one (weight 1), 1 {-1, 1}^D, 2 {-1, 0, 1}^D (weights = likelihood)
};

__global__ void __launch_bounds__(256) k_eval_general(GeneralArgs a) {
    pdl_enter();
    extern __shared__ double gsm[];
    const int G = a.n_groups, NS = a.n_sets, t = a.t, D = a.t + 1;
    double* A = gsm;                    // [G]
    double* Bp = A + G;                 // [G]
    double* Bm = Bp + G;                // [G]
    double* sBp = Bm + G;               // [NS]
    double* sBm = sBp + NS;             // [NS]
    double* red = gsm + ((3 * G + 2 * NS + 1) & ~1);    // [G][8][3] per-warp partials, later [8] (16-byte aligned: the table follows)
    double2* phi_s = reinterpret_cast<double2*>(red + (size_t)G * 24 + 8);      // table of phi_tab
    phi_tab_to_shared(phi_s, a.phi);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_items = *a.count;
    const int64_t N = a.n_nodes;
    const double log_eps = log(kEps), log1p_eps = log(1.0 + kEps);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t i = a.list[item];
        double l[4];
        double s2 = a.v[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            l[j] = 0.0;
            if (j < t) {
                l[j] = a.U[(int64_t)(a.W0 + j) * a.ldu + i];
                s2 = fma(-l[j], l[j], s2);
            }
        }
        const double mi = a.m[i];
        const double s = s2 > 0.0 ? sqrt(s2) : 0.0;
        const double inv_s = s > 0.0 ? 1.0 / s : 0.0;
        const double inv_st = 1.0 / sqrt(fmax(s2, 0.0) + a.noise);
        for (int g = 0; g < G; ++g) {
            double xa = 0.0, xp = 0.0, xm = 0.0;
            for (int q = a.group_begin[g] + threadIdx.x; q < a.group_begin[g + 1]; q += blockDim.x) {
                double num = mi;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < t) num = fma(l[j], a.eta[(int64_t)j * N + q], num);
                const double wq = a.w[q];
                const double cdf = s > 0.0 ? phi_tab(phi_s, num * inv_s) : (num > 0.0 ? 1.0 : 0.0);
                const double dp = (1.0 - num) * inv_st, dm = (-1.0 - num) * inv_st;
                xa = fma(wq, cdf, xa);
                xp = fma(wq, exp(-0.5 * dp * dp), xp);
                xm = fma(wq, exp(-0.5 * dm * dm), xm);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                xa += __shfl_xor_sync(0xffffffffu, xa, o);
                xp += __shfl_xor_sync(0xffffffffu, xp, o);
                xm += __shfl_xor_sync(0xffffffffu, xm, o);
            }
            if (lane == 0) {
                red[(g * 8 + warp) * 3 + 0] = xa;
                red[(g * 8 + warp) * 3 + 1] = xp;
                red[(g * 8 + warp) * 3 + 2] = xm;
            }
        }
        __syncthreads();
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            double xa = 0.0, xp = 0.0, xm = 0.0;
            for (int k = 0; k < 8; ++k) {
                xa += red[(g * 8 + k) * 3 + 0];
                xp += red[(g * 8 + k) * 3 + 1];
                xm += red[(g * 8 + k) * 3 + 2];
            }
            A[g] = xa;
            Bp[g] = xp;
            Bm[g] = xm;
        }
        __syncthreads();
        for (int sidx = threadIdx.x; sidx < NS; sidx += blockDim.x) {
            double xp = 0.0, xm = 0.0;
            for (int g = a.set_group0[sidx]; g < a.set_group0[sidx + 1]; ++g) { xp += Bp[g]; xm += Bm[g]; }
            sBp[sidx] = xp;
            sBm[sidx] = xm;
        }
        __syncthreads();
        const int nr = 1 << D;
        if (a.estimation != 0) {
            // label_estimation = 'optimistic' / 'pessimistic' (ital.py:210-215): no expectation over the relevance
            // configurations -- the largest, resp. the "first or smaller" single term
            //   cur = (log(p'(r | f) + eps) - log(p_r + eps)) * weight(f | r)
            // in the enumeration order of the reference (r over product([F, T]), f over fb_iter; the first sample
            // varies slowest).  The terms are formed in parallel and folded in order by one thread, so that the quirk
            // of the 'pessimistic' fold -- a running value of exactly 0 is replaced by the next term -- is kept.
            double* logp = red + (size_t)G * 24 + 8 + 2 * (size_t)kPhiTableLen;     // [nr]
            double* logq = logp + 32;                                               // [nr][nr]
            double* piece = logq + 1024;                                            // [2^D * |feedback configurations|]
            for (int r = threadIdx.x; r < nr; r += blockDim.x) {
                const int g0 = r & ((1 << t) - 1), rc = r >> t;
                logp[r] = log(fmax(rc ? A[g0] : a.group_mass[g0] - A[g0], 0.0) + kEps);
            }
            for (int pidx = threadIdx.x; pidx < nr * nr; pidx += blockDim.x) {
                const int r = pidx >> D, Om = pidx & (nr - 1), rc = r >> t;
                if (Om == 0) continue;
                const int* e = a.lut + ((size_t)r * nr + Om) * 3;
                double q;
                if (e[2] & 2) q = 1.0;
                else if (e[2] & 1) q = (rc ? Bp[e[0]] : Bm[e[0]]) / fmax(rc ? sBp[e[1]] : sBm[e[1]], 1e-300);
                else q = rc ? A[e[0]] : a.group_mass[e[0]] - A[e[0]];
                logq[pidx] = log(fmin(fmax(q, 0.0), 1.0) + kEps);
            }
            __syncthreads();
            int nf = 1;
            if (a.fb_kind == 1) nf = nr;
            if (a.fb_kind == 2) { nf = 1; for (int j = 0; j < D; ++j) nf *= 3; }
            const int N = nr * nf;
            const double w0 = 1.0 - a.lp, wc = a.lp * (1.0 - a.mp), wm = a.lp * a.mp;
            for (int k = threadIdx.x; k < N; k += blockDim.x) {     // every term of the sequence, in parallel
                const int ri = k / nf;
                int fi = k - ri * nf;
                int rmask = 0, Om = 0;
                bool consistent = true;
                double weight = 1.0;
                // digits of f from the last sample to the first (the first varies slowest); the weight is multiplied
                // up in sample order afterwards
                int fd[5];
                for (int j = D - 1; j >= 0; --j) {
                    const int rj = (ri >> (D - 1 - j)) & 1;
                    rmask |= rj << j;
                    int f;
                    if (a.fb_kind == 0) f = 2 * rj - 1;
                    else if (a.fb_kind == 1) { f = 2 * (fi & 1) - 1; fi >>= 1; }
                    else { f = fi % 3 - 1; fi /= 3; }
                    fd[j] = f;
                    if (f != 0) {
                        Om |= 1 << j;
                        if (f != 2 * rj - 1) consistent = false;
                    }
                }
                if (a.fb_kind != 0) {
                    for (int j = 0; j < D; ++j) {
                        const int rj = (rmask >> j) & 1;
                        weight *= fd[j] == 0 ? w0 : (fd[j] == 2 * rj - 1 ? wc : wm);
                    }
                }
                // nobody labelled: not a feedback configuration (ital.py:201) -- marked NaN and skipped by the fold
                piece[k] = Om == 0 ? nan("") : ((consistent ? logq[rmask * nr + Om] : log_eps) - logp[rmask]) * weight;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                // the fold of ital.py:210-215, literally and in order: 'optimistic' keeps the largest term (from 0),
                // 'pessimistic' replaces a running value of exactly 0 by the next term and otherwise keeps the smaller
                double mi = 0.0;
                for (int k = 0; k < N; ++k) {
                    const double c = piece[k];
                    if (c != c) continue;
                    if (a.estimation == 1) { if (c > mi) mi = c; }
                    else if (mi == 0.0 || c < mi) mi = c;
                }
                a.tags[i] = tag_with_step(a.tags[i], a.epoch, a.t);
                a.score[i] = mi;
                a.gain[i] = mi;
                atomicAdd(a.n_scored, 1);
            }
            __syncthreads();
            continue;
        }
        // assembly over (r, O); O = 0 stands for the -log p_r term
        double part = 0.0;
        for (int pidx = threadIdx.x; pidx < nr * nr; pidx += blockDim.x) {
            const int r = pidx >> D, Om = pidx & (nr - 1);
            const int g0 = r & ((1 << t) - 1), rc = r >> t;
            const double p_r = fmax(rc ? A[g0] : a.group_mass[g0] - A[g0], 0.0);
            double term;
            if (Om == 0) {
                term = -(1.0 - pow(1.0 - a.lp, (double)D)) * log(p_r + kEps);
            } else {
                const int k = __popc(Om);
                const double lam = pow(1.0 - a.lp, (double)(D - k)) * pow(a.lp, (double)k);
                const double c1 = pow(1.0 - a.mp, (double)k);
                const int* e = a.lut + ((size_t)r * nr + Om) * 3;
                double q;
                if (e[2] & 2) q = 1.0;
                else if (e[2] & 1) q = (rc ? Bp[e[0]] : Bm[e[0]]) / fmax(rc ? sBp[e[1]] : sBm[e[1]], 1e-300);
                else q = rc ? A[e[0]] : a.group_mass[e[0]] - A[e[0]];
                q = fmin(fmax(q, 0.0), 1.0);
                term = lam * (c1 * log(q + kEps) + (1.0 - c1) * log_eps);
            }
            part = fma(p_r, term, part);
        }
        (void)log1p_eps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __syncthreads();                                 // red is reused below
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int k = 0; k < 8; ++k) tot += red[k];
            a.tags[i] = tag_with_step(a.tags[i], a.epoch, a.t);
            a.score[i] = tot;
            a.gain[i] = tot;
            atomicAdd(a.n_scored, 1);
        }
        __syncthreads();
    }
}

// ---- shared quadrature nodes on the device (same rule as csrc/snq_host.h / oracle/orthant.py) ----------------
constexpr int kGlStride = 64;            // Gauss-Legendre tables: rule n at [n * 64 .. n * 64 + n)

// Node k of the rule; node index = sum_j digit_j (2q)^(t-1-j).  Every thread recomputes the boundary of each
// dimension from its own prefix of coordinates (a handful of flops) instead of communicating.
template <int T>
__device__ __forceinline__ void snq_node(int64_t k, int64_t N, int q, double R, int q_min, const double* base_m,
                                         const double* base_L, const double* __restrict__ gl_x,
                                         const double* __restrict__ gl_w, double* e, double& wt_out, int& orth_out) {
    const int two_q = 2 * q;
    double wt = 1.0;
    int ob = 0;
    int64_t rem = k, stride = N;
#pragma unroll
    for (int j = 0; j < T; ++j) {
        stride /= two_q;
        const int g = (int)(rem / stride);
        rem -= (int64_t)g * stride;
        double acc = base_m[j];
#pragma unroll
        for (int i = 0; i < T; ++i)
            if (i < j) acc += e[i] * base_L[j * kBaseStride + i];
        const double av = -acc / base_L[j * kBaseStride + j];
        const double c = av < -R ? -R : (av > R ? R : av);
        int n_lo = (int)floor(two_q * (c + R) / (2.0 * R) + 0.5);
        n_lo = max(q_min, min(two_q - q_min, n_lo));
        const int n_hi = two_q - n_lo;
        double x, wg;
        if (g < n_lo) {
            const double half = 0.5 * (c + R);
            x = -R + half * (1.0 + gl_x[n_lo * kGlStride + g]);
            wg = half * gl_w[n_lo * kGlStride + g];
        } else {
            const double half = 0.5 * (R - c);
            x = c + half * (1.0 + gl_x[n_hi * kGlStride + (g - n_lo)]);
            wg = half * gl_w[n_hi * kGlStride + (g - n_lo)];
            ob |= 1 << j;
        }
        wg *= exp(-0.5 * x * x) / 2.50662827463100050242;      // standard normal density
        e[j] = x;
        wt *= wg;
    }
    wt_out = wt;
    orth_out = ob;
}

// ---- change_estimation_subset (ital.py:227-275, AppendedMutualInformation.__call__ ital.py:514-527) --------------
// The batch columns of the running fetch hold ext = [B (tB samples picked so far), S' (the change-estimation subset
// without the members of B)], D = tB + u columns.  The reference integrates over the relevance of B and the candidate
// only and keeps the subset at the signs s* of its means:
//     MI = sum_{r over B + candidate} p_r [ log(P(s*, r | labels of B and the candidate as in r) + eps)
//                                           - log(P(s*, r) + eps) ],     p_r = P(r) with the subset marginalised.
//   p_r and P(s*, r): sums over node sets that depend on ext only (csrc/snq_host.h generate_sub) --
//     part 1, groups [0, G): nodes over B (coordinates of S' zero), candidate variance conditional on B only;
//     part 2, groups [G, 2G): nodes of the prior of ext inside the orthant (r_B, s*).
//   P(s*, r | labels): a labelled sample keeps the sign of its label (variances >> label noise, as in the general
//     feedback model), so this is P(S' keeps s* | labels).  Given the labels of B, S' ~ N(mU_g, CU) and the
//     candidate's label y ~ N(mean_c, tau^2) with covariance c = (Bm Sig) l_i to S': conditioning on y is a rank-one
//     update, S' ~ N(mU_g + c (y - mean_c) / tau^2, CU - c c^T / tau^2), different for every candidate, and its
//     orthant probability is evaluated here with the shared-node rule itself (nodes of the first u - 1 variables
//     generated on the fly by snq_node, the last variable analytic) -- the importance-weighted shortcut over shared
//     nodes (label density times prior nodes) under-resolves candidates that correlate strongly with the subset.
struct SubArgs {
    const int* count;
    const int* list;
    const double* m;
    const double* v;
    const double* U;
    int64_t ldu;
    int W0;
    int tB, D;
    const double* eta;          // dimension-major [D][n_nodes]
    const double* w;
    int64_t n_nodes;
    const int* group_begin;     // 2 G + 1
    const double* mass;         // [2][G]: parts 1 and 2
    const double* mu;           // [G][D]: mean of eta given the labels r_B
    const double* Sig;          // [D][D]: covariance of eta given labels on B
    const double* mU;           // [G][u]: mean of S' given the labels r_B
    const double* CU;           // [u][u]: covariance of S' given labels on B
    const double* BS;           // [u][D]: Bm Sig (covariance of S' with eta)
    int sub_bits;               // s*
    double c1;                  // (1 - mistake_prob)^(tB + 1): probability that no label of the batch + candidate is wrong
    int q_last;                 // Gauss-Legendre nodes per panel of the per-candidate rule (u - 1 variables)
    double R;
    int q_min;
    const double* gl_x;
    const double* gl_w;
    double noise;
    const double2* phi;
    double* score;
    double* gain;
    uint32_t* tags;
    uint32_t epoch;
    int* n_scored;
};

constexpr int kSubMaxCols = 11;          // = kMaxBatch columns of ext
constexpr int kSubMaxU = 5;              // subset members outside the batch

template <int T>
__device__ __forceinline__ double sub_orthant_sum(int64_t N, const SubArgs& a, const double* bm, const double* bL,
                                                  int want, double sgn_last, const double2* phi_s) {
    // sum over the nodes of the first T variables inside orthant `want` of Phi(+-(m_T + L_T. eta) / L_TT)
    double acc = 0.0;
    const double sd = bL[T * kBaseStride + T];
    const double inv_sd = sd > 0.0 ? 1.0 / sd : 0.0;
    for (int64_t k = threadIdx.x; k < N; k += blockDim.x) {
        double e[T > 0 ? T : 1];
        double wt;
        int ob;
        snq_node<T>(k, N, a.q_last, a.R, a.q_min, bm, bL, a.gl_x, a.gl_w, e, wt, ob);
        if (ob != want || wt < 1e-13) continue;         // (the host rule drops the same light nodes)
        double num = bm[T];
#pragma unroll
        for (int j = 0; j < T; ++j) num = fma(bL[T * kBaseStride + j], e[j], num);
        num *= sgn_last;
        const double cdf = sd > 0.0 ? phi_tab(phi_s, num * inv_sd) : (num > 0.0 ? 1.0 : 0.0);
        acc = fma(wt, cdf, acc);
    }
    return acc;
}

__global__ void __launch_bounds__(256) k_eval_sub(SubArgs a) {
    pdl_enter();
    extern __shared__ double ssm[];
    const int G = 1 << a.tB, D = a.D, tB = a.tB, u = a.D - a.tB;
    double* acc = ssm;                                  // [2G]: sums of the groups of parts 1 and 2
    double* qv = ssm + 2 * G;                           // [2G]: P(S' keeps s* | labels), by (g, rc)
    double* red = qv + 2 * G;                           // [8] per-warp partials
    double* bm = red + 8;                               // [8]: mean of S' given all labels
    double* bL = bm + 8;                                // [5][kBaseStride]: Cholesky factor of its covariance
    double2* phi_s = reinterpret_cast<double2*>(ssm + ((4 * G + 16 + 5 * kBaseStride + 1) & ~1));
    phi_tab_to_shared(phi_s, a.phi);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_items = *a.count;
    const int64_t N = a.n_nodes;
    int64_t n_last = 1;
    for (int j = 0; j + 1 < u; ++j) n_last *= 2 * a.q_last;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t i = a.list[item];
        double l[kSubMaxCols];
        double s2B = a.v[i], s2F;
#pragma unroll
        for (int j = 0; j < kSubMaxCols; ++j) l[j] = j < D ? a.U[(int64_t)(a.W0 + j) * a.ldu + i] : 0.0;
        s2F = s2B;
#pragma unroll
        for (int j = 0; j < kSubMaxCols; ++j) {
            if (j < tB) s2B = fma(-l[j], l[j], s2B);
            s2F = fma(-l[j], l[j], s2F);
        }
        const double mi = a.m[i];
        const double sB = s2B > 0.0 ? sqrt(s2B) : 0.0, sF = s2F > 0.0 ? sqrt(s2F) : 0.0;
        const double st2 = fmax(s2F, 0.0) + a.noise;
        // parts 1 and 2: shared nodes
        for (int g = 0; g < 2 * G; ++g) {
            const double sd = g < G ? sB : sF;
            const double inv_sd = sd > 0.0 ? 1.0 / sd : 0.0;
            double x0 = 0.0;
            for (int q = a.group_begin[g] + threadIdx.x; q < a.group_begin[g + 1]; q += blockDim.x) {
                double num = mi;
#pragma unroll
                for (int j = 0; j < kSubMaxCols; ++j)
                    if (j < D) num = fma(l[j], a.eta[(int64_t)j * N + q], num);
                const double cdf = sd > 0.0 ? phi_tab(phi_s, num * inv_sd) : (num > 0.0 ? 1.0 : 0.0);
                x0 = fma(a.w[q], cdf, x0);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x0 += __shfl_xor_sync(0xffffffffu, x0, o);
            if (lane == 0) red[warp] = x0;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int k = 0; k < 8; ++k) tot += red[k];
                acc[g] = tot;
            }
            __syncthreads();
        }
        // part 3: the subset given the labels of B and of the candidate, one rank-one update per (g, label)
        for (int gr = 0; gr < 2 * G; ++gr) {
            const int g = gr & (G - 1), rc = gr >> tB;
            if (u == 0) {
                if (threadIdx.x == 0) qv[gr] = 1.0;
                continue;
            }
            if (threadIdx.x == 0) {
                double lSl = 0.0, mean_c = mi;
                for (int r = 0; r < D; ++r) {
                    double row = 0.0;
                    for (int c = 0; c < D; ++c) row = fma(a.Sig[r * D + c], l[c], row);
                    lSl = fma(l[r], row, lSl);
                    mean_c = fma(l[r], a.mu[g * D + r], mean_c);
                }
                const double tau2 = st2 + fmax(lSl, 0.0);
                const double dy = ((rc ? 1.0 : -1.0) - mean_c) / tau2;
                double cv[kSubMaxU];
                for (int x = 0; x < u; ++x) {
                    double c = 0.0;
                    for (int r = 0; r < D; ++r) c = fma(a.BS[x * D + r], l[r], c);
                    cv[x] = c;
                    bm[x] = fma(c, dy, a.mU[g * u + x]);
                }
                // Cholesky factor of CU - c c^T / tau^2 (pivots floored like the host's)
                for (int x = 0; x < u; ++x)
                    for (int y = 0; y <= x; ++y) {
                        double val = a.CU[x * u + y] - cv[x] * cv[y] / tau2;
                        for (int k = 0; k < y; ++k) val -= bL[x * kBaseStride + k] * bL[y * kBaseStride + k];
                        if (x == y) bL[x * kBaseStride + x] = sqrt(val > 1e-300 ? val : 1e-300);
                        else bL[x * kBaseStride + y] = val / bL[y * kBaseStride + y];
                    }
            }
            __syncthreads();
            const int T = u - 1;
            const int want = a.sub_bits & ((1 << T) - 1);
            const double sgn_last = ((a.sub_bits >> T) & 1) ? 1.0 : -1.0;
            double x0;
            switch (T) {
                case 0: x0 = sub_orthant_sum<0>(n_last, a, bm, bL, want, sgn_last, phi_s); break;
                case 1: x0 = sub_orthant_sum<1>(n_last, a, bm, bL, want, sgn_last, phi_s); break;
                case 2: x0 = sub_orthant_sum<2>(n_last, a, bm, bL, want, sgn_last, phi_s); break;
                case 3: x0 = sub_orthant_sum<3>(n_last, a, bm, bL, want, sgn_last, phi_s); break;
                default: x0 = sub_orthant_sum<4>(n_last, a, bm, bL, want, sgn_last, phi_s); break;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x0 += __shfl_xor_sync(0xffffffffu, x0, o);
            if (lane == 0) red[warp] = x0;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int k = 0; k < 8; ++k) tot += red[k];
                qv[gr] = tot;
            }
            __syncthreads();
        }
        __syncthreads();
        // assembly over the 2 G relevance configurations of B + candidate
        double term = 0.0;
        if (threadIdx.x < 2 * G) {
            const int g = threadIdx.x & (G - 1), rc = threadIdx.x >> tB;
            const double p_r = fmax(rc ? acc[g] : a.mass[g] - acc[g], 0.0);
            const double P = fmax(rc ? acc[G + g] : a.mass[G + g] - acc[G + g], 0.0);
            const double q = fmin(fmax(qv[threadIdx.x], 0.0), 1.0);
            // a user who mislabels (ital.py:317-328): any wrong label contradicts r and leaves probability ~0 for it
            term = p_r * (a.c1 * log(q + kEps) + (1.0 - a.c1) * log(kEps) - log(P + kEps));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
        if (lane == 0) red[warp] = term;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int k = 0; k < 8; ++k) tot += red[k];
            a.tags[i] = tag_with_step(a.tags[i], a.epoch, a.D);
            a.score[i] = tot;
            a.gain[i] = tot;
            atomicAdd(a.n_scored, 1);
        }
        __syncthreads();
    }
}

// ---- clip_cov: grouped orthant probabilities for more than 5 samples (ital.py:360-362, 386-429, 590-616) -----------
// From the sixth sample of a batch on, the reference drops correlations below clip_cov and multiplies the orthant
// probabilities of the resulting independent groups.  For users who label everything the score is the entropy of the
// sign pattern, which is additive over independent groups: score(i) = sum over the groups of the batch that candidate i
// is not connected to of H(group) + H({i} + the groups it connects).  k_clip_adj finds, per candidate, the union S of
// the batch's groups it is connected to (|corr| > clip_cov with any member); candidates with S empty are finished in
// closed form, the others are listed for k_eval_clip, which scores them with the node set of S (generated on the host
// for every S that occurs, csrc/ital_capi.cu propose_clip).
struct ClipDesc {
    long long eta_off;          // node coordinates of the set, dimension-major [dims][n]
    int n, dims;
    int gb_off;                 // 2^dims + 1 orthant offsets (relative to the set)
    int mass_off;               // 2^dims orthant masses
    int T_off;                  // [dims][t]: projection of a candidate on the set's own Cholesky factor
    int pad;                    // offset of the set's weights
    double h_rest;              // entropy of the groups outside S
};

struct ClipArgs {
    const int* count;           // candidates (k_worklist, all of them)
    const int* list;
    int* count2;                // candidates connected to the batch
    int* list2;
    const double* m;
    const double* v;
    const double* U;
    int64_t ldu;
    int W0, t;
    const double* base_L;       // [t][kBaseStride]
    const double* sd_b;         // [t] standard deviations of the batch members
    const int* comp_mask;       // [t] members of the group of batch member b, as a bit mask
    double th, h_all;           // clip_cov; entropy of all groups of the batch
    unsigned short* sub;        // [n] S per candidate
    unsigned* present;          // [32] which S occur
    const ClipDesc* desc;       // [1024] by S
    const double* eta;
    const double* w;
    const int* gb;
    const double* mass;
    const double* T;
    const double2* phi;
    double log1p_eps;
    double* score;
    double* gain;
    uint32_t* tags;
    uint32_t epoch;
    int* n_scored;
};

__global__ void __launch_bounds__(256) k_clip_adj(ClipArgs a) {
    pdl_enter();
    const int n_items = *a.count;
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
        const int64_t i = a.list[item];
        const double vi = a.v[i];
        double l[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) l[j] = j < a.t ? a.U[(int64_t)(a.W0 + j) * a.ldu + i] : 0.0;
        int S = 0;
        if (vi > 0.0) {
            const double sdi = sqrt(vi);
            for (int b = 0; b < a.t; ++b) {
                double cov = 0.0;
                for (int j = 0; j <= b; ++j) cov = fma(a.base_L[b * kBaseStride + j], l[j], cov);
                if (fabs(cov / (a.sd_b[b] * sdi)) > a.th) S |= a.comp_mask[b];
            }
        }
        a.sub[i] = (unsigned short)S;
        if (S == 0) {
            // on its own: two-sided entropy of its sign, plus the groups of the batch
            const double sd = vi > 0.0 ? sqrt(vi) : 0.0;
            const double mi = a.m[i];
            const double p1 = sd > 0.0 ? 0.5 * erfc(-mi / sd * 0.70710678118654752440) : (mi > 0.0 ? 1.0 : 0.0);
            const double sc = mi_term(p1, a.log1p_eps) + mi_term(1.0 - p1, a.log1p_eps) + a.h_all;
            a.tags[i] = tag_with_step(a.tags[i], a.epoch, a.t);
            a.score[i] = sc;
            a.gain[i] = sc;
            atomicAdd(a.n_scored, 1);
        } else {
            atomicOr(a.present + (S >> 5), 1u << (S & 31));
            a.list2[atomicAdd(a.count2, 1)] = (int)i;
        }
    }
}

__global__ void __launch_bounds__(256) k_eval_clip(ClipArgs a) {
    pdl_enter();
    __shared__ double red[8];
    __shared__ double xs[10];
    __shared__ double2 phi_s[kPhiTableLen];
    phi_tab_to_shared(phi_s, a.phi);
    __syncthreads();
    const int n_items = *a.count2;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t i = a.list2[item];
        const ClipDesc dsc = a.desc[a.sub[i]];
        const int dims = dsc.dims;
        __syncthreads();
        if ((int)threadIdx.x < dims) {                  // projection on the set's own factor: x = T l_i
            double acc = 0.0;
            for (int c = 0; c < a.t; ++c)
                acc = fma(a.T[dsc.T_off + threadIdx.x * a.t + c], a.U[(int64_t)(a.W0 + c) * a.ldu + i], acc);
            xs[threadIdx.x] = acc;
        }
        __syncthreads();
        double x[10];
        double s2 = a.v[i];
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            x[j] = j < dims ? xs[j] : 0.0;
            s2 = fma(-x[j], x[j], s2);
        }
        const double mi = a.m[i];
        const double sd = s2 > 0.0 ? sqrt(s2) : 0.0;
        const double inv_s = sd > 0.0 ? 1.0 / sd : 0.0;
        const double* eta = a.eta + dsc.eta_off;
        const double* w = a.w + dsc.pad;
        const int* gb = a.gb + dsc.gb_off;
        const double* mass = a.mass + dsc.mass_off;
        const int64_t N = dsc.n;
        double sc = 0.0;
        const int nb = 1 << dims;
        for (int b = 0; b < nb; ++b) {
            double acc = 0.0;
            for (int q = gb[b] + threadIdx.x; q < gb[b + 1]; q += 256) {
                double num = mi;
#pragma unroll
                for (int j = 0; j < 10; ++j)
                    if (j < dims) num = fma(x[j], eta[(int64_t)j * N + q], num);
                const double cdf = sd > 0.0 ? phi_tab(phi_s, num * inv_s) : (num > 0.0 ? 1.0 : 0.0);
                acc = fma(w[q], cdf, acc);
            }
            const double p_plus = team_sum(acc, 256, red);
            const double p_minus = fmax(mass[b] - p_plus, 0.0);
            sc += mi_term(p_plus, a.log1p_eps) + mi_term(p_minus, a.log1p_eps);
        }
        if (threadIdx.x == 0) {
            sc += dsc.h_rest;
            a.tags[i] = tag_with_step(a.tags[i], a.epoch, a.t);
            a.score[i] = sc;
            a.gain[i] = sc;
            atomicAdd(a.n_scored, 1);
        }
    }
}

// ---- tensor rule for 4 and 5 base variables on the device ------------------------------------------------------------
// 331 776 / 3.2 million nodes are generated twice (a node costs a few hundred flops, storing them all would cost more):
// once to count the kept nodes per orthant and block, once to scatter them to their place in the orthant-sorted,
// dimension-major list -- generation order within an orthant, like snq_host.h generate().
template <int T>
__global__ void __launch_bounds__(256) k_tn_count(int q, double R, int q_min, const double* __restrict__ base_m,
                                                  const double* __restrict__ base_L, const double* __restrict__ gl_x,
                                                  const double* __restrict__ gl_w, int64_t N, double w_min,
                                                  int* __restrict__ blk_cnt) {
    pdl_enter();
    constexpr int NB = 1 << T;
    __shared__ int cnt_s[NB];
    if (threadIdx.x < NB) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k < N) {
        double e[T], wt;
        int ob;
        snq_node<T>(k, N, q, R, q_min, base_m, base_L, gl_x, gl_w, e, wt, ob);
        if (wt >= w_min) atomicAdd(&cnt_s[ob], 1);
    }
    __syncthreads();
    if (threadIdx.x < NB) blk_cnt[(int64_t)blockIdx.x * NB + threadIdx.x] = cnt_s[threadIdx.x];
}

// exclusive scan of the block counts per orthant (a warp per orthant), orthant offsets
__global__ void __launch_bounds__(1024) k_tn_scan(int nb, int n_blocks, const int* __restrict__ blk_cnt,
                                                  int* __restrict__ blk_off, int* __restrict__ group_begin,
                                                  int* __restrict__ n_kept) {
    pdl_enter();
    __shared__ int tot[32];
    const int lane = threadIdx.x & 31, o = threadIdx.x >> 5;
    if (o < nb) {
        int run = 0;
        for (int b0 = 0; b0 < n_blocks; b0 += 32) {
            const int b = b0 + lane;
            const int c = b < n_blocks ? blk_cnt[(int64_t)b * nb + o] : 0;
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += v;
            }
            if (b < n_blocks) blk_off[(int64_t)b * nb + o] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) tot[o] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int pos = 0;
        for (int b = 0; b < nb; ++b) { group_begin[b] = pos; pos += tot[b]; }
        group_begin[nb] = pos;
        *n_kept = pos;
    }
}

template <int T>
__global__ void __launch_bounds__(256) k_tn_scatter(int q, double R, int q_min, const double* __restrict__ base_m,
                                                    const double* __restrict__ base_L, const double* __restrict__ gl_x,
                                                    const double* __restrict__ gl_w, int64_t N, double w_min,
                                                    const int* __restrict__ blk_off, const int* __restrict__ group_begin,
                                                    int64_t stride, double* __restrict__ eta, double* __restrict__ w) {
    pdl_enter();
    constexpr int NB = 1 << T;
    __shared__ int warp_cnt[8][NB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < 8 * NB; k += 256) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double e[T], wt = 0.0;
    int ob = 0;
    bool keep = false;
    if (k < N) {
        snq_node<T>(k, N, q, R, q_min, base_m, base_L, gl_x, gl_w, e, wt, ob);
        keep = wt >= w_min;
    }
    // rank among the kept nodes of the same orthant with a lower index: within the warp, then over the warps before
    const unsigned kept_mask = __ballot_sync(0xffffffffu, keep);
    int rank = 0;
    if (keep) {
        const unsigned same = __match_any_sync(kept_mask, ob);
        rank = __popc(same & ((1u << lane) - 1u));
        if (rank == 0) warp_cnt[warp][ob] = __popc(same);
    }
    __syncthreads();
    if (keep) {
        int before = 0;
        for (int ww = 0; ww < warp; ++ww) before += warp_cnt[ww][ob];
        const int64_t pos = (int64_t)group_begin[ob] + blk_off[(int64_t)blockIdx.x * NB + ob] + before + rank;
#pragma unroll
        for (int j = 0; j < T; ++j) eta[(int64_t)j * stride + pos] = e[j];
        w[pos] = wt;
    }
}

// orthant masses (one block per orthant: fixed strided partial sums, fixed tree), then H(base) and the total mass
__global__ void __launch_bounds__(256) k_tn_masses(const int* __restrict__ group_begin, const double* __restrict__ w,
                                                   double* __restrict__ masses) {
    pdl_enter();
    __shared__ double red[8];
    const int g0 = group_begin[blockIdx.x], g1 = group_begin[blockIdx.x + 1];
    double sum = 0.0;
    for (int k = g0 + threadIdx.x; k < g1; k += 256) sum += w[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < 8; ++k) tot += red[k];
        masses[blockIdx.x] = tot;
    }
}

__global__ void k_tn_hbase(int nb, const double* __restrict__ masses, double log1p_eps, double* __restrict__ hbase) {
    pdl_enter();
    if (threadIdx.x == 0) {
        double h = 0.0, tot = 0.0;
        for (int b = 0; b < nb; ++b) {
            h += masses[b] * (log1p_eps - log(masses[b] + kEps));
            tot += masses[b];
        }
        hbase[0] = h;
        hbase[1] = tot;
    }
}

// ---- sequential-conditioning lattice on the device (t >= 6 base variables; same rule as snq_host.h generate_sc) ----
// inverse of the standard normal CDF: Abramowitz-Stegun 26.2.23 start, Halley steps on 0.5 erfc(-x / sqrt 2)
__device__ __forceinline__ double ndtri_dev(double p) {
    const bool lower = p < 0.5;
    const double pp = lower ? p : 1.0 - p;
    const double tt = sqrt(-2.0 * log(pp));
    double x = tt - (2.515517 + 0.802853 * tt + 0.010328 * tt * tt) /
                        (1.0 + 1.432788 * tt + 0.189269 * tt * tt + 0.001308 * tt * tt * tt);
    x = lower ? -x : x;
    for (int it = 0; it < 6; ++it) {
        const double cdf = 0.5 * erfc(-x * 0.70710678118654752440);
        const double pdf = exp(-0.5 * x * x) * 0.39894228040143267794;
        const double f = cdf - p;
        const double dx = f / (pdf + 0.5 * x * f);
        x -= dx;
        if (fabs(dx) < 1e-15 * (1.0 + fabs(x))) break;
    }
    return x;
}

struct ScArgs {
    int t;
    double alpha[12];               // fractional parts of the square roots of the first primes (Kronecker sequence)
    const double* base_m;
    const double* base_L;           // [t][kBaseStride]
    int64_t n_total;                // node budget over all orthants (kScN)
    int pilot, n_min;               // kScPilot, kScMin
    double p_min;                   // kScPMin
    int64_t stride;                 // of the dimension-major coordinates (capacity)
    double* P;                      // [2^t] pilot masses
    int* cnt;                       // [2^t]
    int* group_begin;               // [2^t + 1]
    int* chunk_orth;                // per chunk of up to 256 nodes: orthant, offset within the orthant
    int* chunk_off;
    int* n_chunks;
    double* chunk_sum;
    double* eta;
    double* w;
    double* masses;
    double* hbase;                  // {H(base), total mass}
    double log1p_eps;
};

// node k of N inside orthant b: coordinates into e[], returns the weight
__device__ __forceinline__ double sc_node(const ScArgs& a, const double* bm, const double* bL, int b, int64_t k,
                                          int64_t N, double* e) {
    double wk = 1.0 / (double)N;
    for (int j = 0; j < a.t; ++j) {
        double u = ((double)k + 0.5) * a.alpha[j];
        u -= floor(u);
        u = 1.0 - fabs(2.0 * u - 1.0);
        double acc = bm[j];
        for (int i = 0; i < j; ++i) acc += bL[j * kBaseStride + i] * e[i];
        const double av = -acc / bL[j * kBaseStride + j];
        if ((b >> j) & 1) {                             // z_j > 0: eta_j above the boundary
            const double q = 0.5 * erfc(av * 0.70710678118654752440);
            const double p = u * q;
            e[j] = -ndtri_dev(p < 1e-300 ? 1e-300 : p);
            wk *= q;
        } else {
            const double q = 0.5 * erfc(-av * 0.70710678118654752440);
            const double p = u * q;
            e[j] = ndtri_dev(p < 1e-300 ? 1e-300 : p);
            wk *= q;
        }
    }
    return wk;
}

__device__ __forceinline__ double block_sum_fixed(double x, double* red) {     // 256 threads, fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += red[k];
    return tot;
}

// pilot pass: one block per orthant, one pilot node per thread
__global__ void __launch_bounds__(256) k_sc_pilot(ScArgs a) {
    pdl_enter();
    __shared__ double red[8];
    __shared__ double bm[16], bL[16 * kBaseStride];
    for (int k = threadIdx.x; k < 16 * kBaseStride; k += 256) bL[k] = a.base_L[k];
    if (threadIdx.x < 16) bm[threadIdx.x] = a.base_m[threadIdx.x];
    __syncthreads();
    double e[12];
    const double wk = (int)threadIdx.x < a.pilot ? sc_node(a, bm, bL, blockIdx.x, threadIdx.x, a.pilot, e) : 0.0;
    const double tot = block_sum_fixed(wk, red);
    if (threadIdx.x == 0) a.P[blockIdx.x] = tot;
}

// node counts per orthant in proportion to the pilot masses, offsets, list of chunks (one block)
__global__ void __launch_bounds__(1024) k_sc_alloc(ScArgs a) {
    pdl_enter();
    __shared__ double s_total;
    __shared__ int s_cnt[1024];
    const int nb = 1 << a.t;
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int b = 0; b < nb; ++b) total += a.P[b];
        s_total = total;
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) {
        const double Pb = a.P[threadIdx.x];
        int c = 0;
        if (Pb >= a.p_min) {
            c = (int)floor((double)a.n_total * Pb / s_total + 0.5);
            if (c < a.n_min) c = a.n_min;
        }
        s_cnt[threadIdx.x] = c;
        a.cnt[threadIdx.x] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int pos = 0, nc = 0;
        for (int b = 0; b < nb; ++b) {
            a.group_begin[b] = pos;
            for (int off = 0; off < s_cnt[b]; off += 256) {
                a.chunk_orth[nc] = b;
                a.chunk_off[nc] = off;
                ++nc;
            }
            pos += s_cnt[b];
        }
        a.group_begin[nb] = pos;
        *a.n_chunks = nc;
    }
}

// the nodes: one block per chunk of up to 256 nodes of one orthant
__global__ void __launch_bounds__(256) k_sc_generate(ScArgs a) {
    pdl_enter();
    __shared__ double red[8];
    __shared__ double bm[16], bL[16 * kBaseStride];
    for (int k = threadIdx.x; k < 16 * kBaseStride; k += 256) bL[k] = a.base_L[k];
    if (threadIdx.x < 16) bm[threadIdx.x] = a.base_m[threadIdx.x];
    __syncthreads();
    const int nc = *a.n_chunks;
    for (int c = blockIdx.x; c < nc; c += gridDim.x) {
        const int b = a.chunk_orth[c], off = a.chunk_off[c];
        const int N = a.cnt[b];
        const int k = off + threadIdx.x;
        double wk = 0.0;
        if (k < N) {
            double e[12];
            wk = sc_node(a, bm, bL, b, k, N, e);
            const int64_t pos = a.group_begin[b] + k;
            for (int j = 0; j < a.t; ++j) a.eta[(int64_t)j * a.stride + pos] = e[j];
            a.w[pos] = wk;
        }
        const double tot = block_sum_fixed(wk, red);
        if (threadIdx.x == 0) a.chunk_sum[c] = tot;
    }
}

// orthant masses from the chunk sums (chunk order), total, H(base) of the normalised masses (one block)
__global__ void __launch_bounds__(1024) k_sc_masses(ScArgs a) {
    pdl_enter();
    __shared__ double s_m[1024];
    __shared__ double s_tot;
    const int nb = 1 << a.t, nc = *a.n_chunks;
    if ((int)threadIdx.x < nb) {
        // the chunks of an orthant are consecutive in the chunk list: find the first by the offsets
        double sum = 0.0;
        for (int c = 0; c < nc; ++c)
            if (a.chunk_orth[c] == (int)threadIdx.x) sum += a.chunk_sum[c];
        s_m[threadIdx.x] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int b = 0; b < nb; ++b) tot += s_m[b];
        s_tot = tot;
    }
    __syncthreads();
    if ((int)threadIdx.x < nb) {
        s_m[threadIdx.x] /= s_tot;
        a.masses[threadIdx.x] = s_m[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double h = 0.0, tot = 0.0;
        for (int b = 0; b < nb; ++b) {
            h += s_m[b] * (a.log1p_eps - log(s_m[b] + kEps));
            tot += s_m[b];
        }
        a.hbase[0] = h;
        a.hbase[1] = tot;
        a.P[0] = s_tot;                                 // (the scale of the weights, for k_sc_scale)
    }
}

// the orthant masses must add up to one: scaling the weights accordingly removes the error all node sets share
__global__ void __launch_bounds__(256) k_sc_scale(ScArgs a) {
    pdl_enter();
    const int n = a.group_begin[1 << a.t];
    const double inv = a.P[0];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) a.w[k] /= inv;
}

template <int T>
__global__ void __launch_bounds__(256) k_snq_generate(int q, double R, int q_min, const double* __restrict__ base_m,
                                                      const double* __restrict__ base_L,
                                                      const double* __restrict__ gl_x, const double* __restrict__ gl_w,
                                                      int64_t N, double* __restrict__ eta, double* __restrict__ w,
                                                      int* __restrict__ orth) {
    pdl_enter();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    double e[T];
    double wt;
    int ob;
    snq_node<T>(k, N, q, R, q_min, base_m, base_L, gl_x, gl_w, e, wt, ob);
#pragma unroll
    for (int j = 0; j < T; ++j) eta[(int64_t)j * N + k] = e[j];
    w[k] = wt;
    orth[k] = ob;
}

// Base orthant probabilities P_b = sum of the kept weights per orthant, the score of the base alone and the total
// mass, in a fixed two-level order shared by the one-CTA finalize kernel and the persistent fetch kernel (where every
// CTA generates one chunk of the nodes): the nodes are cut into `n_chunks` chunks of ceil(N / n_chunks) consecutive
// generated nodes; inside a chunk lane l of a warp adds the kept nodes among the generated nodes l, l + 32, ... of the
// chunk and the lanes are combined by an xor butterfly (chunk_mass_warp); the chunk sums are then added in ascending chunk order
// (masses_from_chunks).  t <= 3.
__device__ __forceinline__ void chunk_mass_warp(const double* w, const int* orth, int cnt, int lane, double w_min,
                                                double* out8) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = lane; k < cnt; k += 32) {
        const int ob = orth[k];
        const double wk = w[k];
        if (wk >= w_min) {
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[b] += (ob == b) ? wk : 0.0;
        }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
        if (lane == 0) out8[b] = acc[b];
    }
}

// chunk_part: [n_chunks][8] in shared memory.  Every thread of the CTA must call it (block barriers inside); thread b < 8
// adds orthant b over the chunks, thread 0 writes h_out[0] = H(base), h_out[1] = total mass.
__device__ __forceinline__ void masses_from_chunks(int t, int n_chunks, const double* chunk_part, double log1p_eps,
                                                   double* masses, double* h_out) {
    if (threadIdx.x < 8) {
        double p = 0.0;
        for (int c = 0; c < n_chunks; ++c) p += chunk_part[c * 8 + threadIdx.x];
        masses[threadIdx.x] = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nb = 1 << t;
        double h = 0.0, tot = 0.0;
        for (int b = 0; b < nb; ++b) {
            const double p = masses[b];
            h += mi_term(p, log1p_eps);
            tot += p;
        }
        h_out[0] = h;
        h_out[1] = tot;                                 // total mass (1 up to quadrature error)
    }
    __syncthreads();
}

// Drops the nodes whose weight is below w_min (about half of the nodes at t = 3 carry a total mass of ~1e-11), sorts
// the kept ones by orthant (stable: generation order inside an orthant) and packs them as {eta_0, eta_1, eta_2,
// weight} records, every orthant zero-padded to a multiple of kNodePad nodes (group_begin[2^t + 1] = offsets); then
// sums the base orthant probabilities and the score of the base alone (chunk_mass_warp / masses_from_chunks).  One
// block; every warp owns a contiguous range of generated nodes and walks it 32 at a time (coalesced); ballots per
// orthant give the stable ranks inside the warp, the warp totals the offsets between warps.
// N <= 13824 = 32 warps x kIt x 32 nodes.
__global__ void __launch_bounds__(1024) k_snq_finalize(int t, int64_t N, double w_min,
                                                       const double* __restrict__ eta_in, const double* __restrict__ w_in,
                                                       const int* __restrict__ orth_in, double* __restrict__ nodes4,
                                                       int* __restrict__ group_begin,
                                                       double log1p_eps, double* __restrict__ masses,
                                                       double* __restrict__ h_base, int* __restrict__ n_kept,
                                                       int n_chunks) {
    pdl_enter();
    extern __shared__ double fin_sm[];                  // [n_chunks][8] chunk sums, [8] masses
    __shared__ int warp_tot[32][8];
    __shared__ int gb[9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nb = 1 << t;
    constexpr int kIt = 14;
    const int64_t per = ((N + nwarps - 1) / nwarps + 31) / 32 * 32;
    const int64_t k0 = min(N, (int64_t)warp * per), k1 = min(N, k0 + per);
    const unsigned lt = (1u << lane) - 1u;
    double wv[kIt];
    int ov[kIt], rk[kIt];                               // orthant and rank inside (warp, orthant); -1 = dropped
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int64_t k = k0 + it * 32 + lane;
        wv[it] = k < k1 ? w_in[k] : 0.0;
        ov[it] = k < k1 ? orth_in[k] : 0;
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const bool keep = wv[it] >= w_min && k0 + it * 32 + lane < k1;
        rk[it] = -1;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const unsigned bal = __ballot_sync(0xffffffffu, keep && ov[it] == o);
            if (keep && ov[it] == o) rk[it] = cnt[o] + __popc(bal & lt);
            cnt[o] += __popc(bal);
        }
    }
    if (lane < 8) {
        int c = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) c = (lane == o) ? cnt[o] : c;
        warp_tot[warp][lane] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0, kept = 0;
        for (int o = 0; o < nb; ++o) {
            int all = 0;
            for (int ww = 0; ww < nwarps; ++ww) all += warp_tot[ww][o];
            gb[o] = g;
            g += pad_nodes(all);
            kept += all;
        }
        gb[nb] = g;
        *n_kept = kept;
    }
    __syncthreads();
    if (threadIdx.x <= nb) group_begin[threadIdx.x] = gb[threadIdx.x];
    int off[8];                                         // where this warp's nodes of every orthant start
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        int before = 0, all = 0;
        for (int ww = 0; ww < nwarps; ++ww) {
            const int c = warp_tot[ww][o];
            if (ww < warp) before += c;
            all += c;
        }
        off[o] = (o < nb ? gb[o] : 0) + before;
        // zero-weight tail of the orthant (all .. padded), written by the warp with the same number
        if (o < nb && warp == o)
            for (int k = gb[o] + all + lane; k < gb[o + 1]; k += 32) {
                nodes4[4 * (size_t)k + 0] = 0.0;
                nodes4[4 * (size_t)k + 1] = 0.0;
                nodes4[4 * (size_t)k + 2] = 0.0;
                nodes4[4 * (size_t)k + 3] = 0.0;
            }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        if (rk[it] < 0) continue;
        int base = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) base = (ov[it] == o) ? off[o] : base;
        const size_t pos = (size_t)(base + rk[it]);
        const int64_t k = k0 + it * 32 + lane;
        nodes4[4 * pos + 0] = eta_in[k];
        nodes4[4 * pos + 1] = t >= 2 ? eta_in[N + k] : 0.0;
        nodes4[4 * pos + 2] = t >= 3 ? eta_in[2 * N + k] : 0.0;
        nodes4[4 * pos + 3] = wv[it];
    }
    // masses in the two-level order of the persistent fetch kernel: chunk c = generated nodes [c per, (c + 1) per)
    double* chunk_part = fin_sm;
    double* sm_masses = fin_sm + (size_t)n_chunks * 8;
    const int64_t per_chunk = (N + n_chunks - 1) / n_chunks;
    for (int c = warp; c < n_chunks; c += nwarps) {
        const int64_t c0 = min(N, (int64_t)c * per_chunk), c1 = min(N, c0 + per_chunk);
        chunk_mass_warp(w_in + c0, orth_in + c0, (int)(c1 - c0), lane, w_min, chunk_part + c * 8);
    }
    __syncthreads();
    masses_from_chunks(t, n_chunks, chunk_part, log1p_eps, sm_masses, h_base);      // t <= 3 here
    if (threadIdx.x < nb) masses[threadIdx.x] = sm_masses[threadIdx.x];
}

// np.argmax over the shards' proposals + AppendedMutualInformation.append (ital/ital.py:130-131, 561-568): choose
// the best of G point records (score desc, global row asc), make it the record k_extend will read, and append it
// to the batch state kept on the device (mean, Cholesky row of the batch's posterior covariance, selection list).
struct PeerWait {                        // peer-memory exchange (peer_put in k_record): wait until every shard's record of this
    const unsigned long long* flags = nullptr;   // epoch has landed in local memory; nullptr = records already there
    unsigned long long epoch = 0;
    int* error = nullptr;                // set to 1 if a peer did not deliver within the time limit
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) k_pick_winner(const double* recs, int G, int64_t rec_stride, int64_t rec_len, int t,
                                                     int W, double* __restrict__ rec_in, double* __restrict__ base_m,
                                                     double* __restrict__ base_L, double* __restrict__ sel,
                                                     double* __restrict__ rec_hist, uint8_t* __restrict__ mask,
                                                     int64_t row_offset, int64_t n, uint8_t mark_bits, PeerWait pw) {
    pdl_enter();
    __shared__ int win_s;
    if (pw.flags != nullptr) {
        // records were stored into this GPU's memory by the peers (peer_put in their k_record, over NVLink): spin on their epoch flags,
        // bounded (a peer that never delivers must not hang the GPU), and read the records past L1 (__ldcg)
        if ((int)threadIdx.x < G) {
            const unsigned long long t0 = global_ns();
            unsigned spins = 0;
            while (ld_acquire_sys(pw.flags + threadIdx.x) < pw.epoch) {
                if ((++spins & 1023u) == 0 && global_ns() - t0 > 5000000000ull) { *pw.error = 1; break; }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int win = -1;
        for (int g = 0; g < G; ++g) {
            const double idx = __ldcg(recs + g * rec_stride), sc = __ldcg(recs + g * rec_stride + 1);
            if (idx < 0.0 || sc != sc) continue;
            if (win < 0 || sc > __ldcg(recs + win * rec_stride + 1) ||
                (sc == __ldcg(recs + win * rec_stride + 1) && idx < __ldcg(recs + win * rec_stride)))
                win = g;
        }
        win_s = win;
    }
    __syncthreads();
    const int win = win_s;
    if (win < 0) {
        if (threadIdx.x == 0) {
            rec_in[0] = -1.0;
            rec_in[1] = -INFINITY;
            sel[2 * t] = -1.0;
            sel[2 * t + 1] = -INFINITY;
        }
        return;
    }
    const double* r = recs + (int64_t)win * rec_stride;
    for (int64_t k = threadIdx.x; k < rec_len; k += blockDim.x) {
        const double val = __ldcg(r + k);
        rec_in[k] = val;
        rec_hist[(int64_t)t * rec_len + k] = val;       // kept for rows that catch up later (k_catchup)
    }
    if (threadIdx.x == 0) {
        const long long loc = (long long)__ldcg(r) - row_offset;
        if (loc >= 0 && loc < n) mask[loc] |= mark_bits;    // the chosen row leaves the candidate set
        base_m[t] = __ldcg(r + 2);
        for (int j = 0; j < t; ++j) base_L[t * kBaseStride + j] = __ldcg(r + 8 + W + j);
        base_L[t * kBaseStride + t] = sqrt(fmax(__ldcg(r + 3), 1e-300));
        sel[2 * t] = __ldcg(r);
        sel[2 * t + 1] = __ldcg(r + 1);
    }
}

// The all-gather of a greedy step without NCCL: the kernel that writes a shard's proposal (k_record) also stores it
// into the slot reserved for this shard in every shard's exchange buffer (peer memory mapped through CUDA IPC; the
// stores travel over NVLink / NVSwitch), then publishes the epoch in the flag words with system-scope release
// semantics.  Slots are double-buffered by epoch parity: a shard can run at most one step ahead of a peer that still
// reads the previous record.
//   buffer of a shard: [G flag words, padded to 256 bytes][2][G][slot_doubles]
struct PeerPut {
    unsigned char* const* peer_base = nullptr;   // device array of G mapped buffers; nullptr = no exchange
    int G = 0;
    int rank = 0;
    int64_t slot_doubles = 0;
    unsigned long long epoch = 0;
};

__device__ __forceinline__ void peer_put(const PeerPut& pp, const double* rec, int64_t rec_len) {
    __syncthreads();                                     // the record is complete (written by this block)
    for (int g = 0; g < pp.G; ++g) {
        double* dst = reinterpret_cast<double*>(pp.peer_base[g] + 256) +
                      ((int64_t)(pp.epoch & 1) * pp.G + pp.rank) * pp.slot_doubles;
        for (int64_t k = threadIdx.x; k < rec_len; k += blockDim.x) dst[k] = rec[k];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pp.G) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(pp.peer_base[threadIdx.x]) + pp.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(pp.epoch) : "memory");
    }
}

__global__ void k_fill(double* __restrict__ p, int64_t n, double value) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = value;
}

// mask[i] = (mask[i] & ~clear) | set for the listed local rows
__global__ void k_mask_rows(uint8_t* __restrict__ mask, const int64_t* __restrict__ rows, int64_t m,
                            uint8_t set_bits) {
    pdl_enter();
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x)
        mask[rows[k]] |= set_bits;
}

__global__ void k_mask_all(uint8_t* __restrict__ mask, int64_t n, uint8_t and_bits, uint8_t or_bits) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mask[i] = (mask[i] & and_bits) | or_bits;
}

__global__ void k_mask_clear_rows(uint8_t* __restrict__ mask, const int64_t* __restrict__ rows, int64_t m,
                                  uint8_t clear_bits) {
    pdl_enter();
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x)
        mask[rows[k]] &= (uint8_t)~clear_bits;
}

// Point record of local row `row` (or of best->idx when row < 0); see ITAL_RECORD_HEADER in ital_b200.h.
template <typename XT>
__global__ void __launch_bounds__(256) k_record(long long row, const Best* __restrict__ best, int64_t row_offset,
                                                const XT* __restrict__ X, int d, int d_pad,
                                                const double* __restrict__ sqn, const double* __restrict__ m,
                                                const double* __restrict__ v, const double* __restrict__ U,
                                                int64_t ldu, int W_lab, int W_tot, int w_cap,
                                                const double* __restrict__ gain, double* __restrict__ rec,
                                                double shift_coef, const double* __restrict__ h_base,
                                                int* __restrict__ counters = nullptr,
                                                int* __restrict__ counters_dst = nullptr,
                                                CommitTargets ct = CommitTargets(), PickSrc src = PickSrc(),
                                                PeerPut pp = PeerPut()) {
    pdl_enter();
    __shared__ Best pick_s;
    if (row < 0 && (src.block_best != nullptr || src.list != nullptr)) {
        // the argmax that used to be a kernel of its own (a final one-block reduction after k_score0 / k_argmax_list)
        double bs = 0.0;
        long long bi = -1;
        if (src.block_best != nullptr) {
            for (int k = threadIdx.x; k < src.nblocks; k += blockDim.x)
                if (better(src.block_best[k].score, src.block_best[k].idx, bs, bi)) {
                    bs = src.block_best[k].score;
                    bi = src.block_best[k].idx;
                }
        } else {
            const int cnt = *src.count;
            for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
                const long long i = src.list[k];
                if (better(src.score[i], i, bs, bi)) { bs = src.score[i]; bi = i; }
            }
        }
        __shared__ double pss[8];
        __shared__ long long psi[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { pss[threadIdx.x >> 5] = bs; psi[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
                if (better(pss[w], psi[w], bs, bi)) { bs = pss[w]; bi = psi[w]; }
            pick_s.score = bs;
            pick_s.idx = bi;
        }
        __syncthreads();
        best = &pick_s;
    }
    if (counters_dst != nullptr && threadIdx.x < 4) {
        counters_dst[threadIdx.x] = counters[threadIdx.x];
        if (threadIdx.x < 3) counters[threadIdx.x] = 0;     // ready for the next greedy step (no memset in between)
    }
    // shift_coef * (total mass): what a user who mislabels with probability mistake_prob adds to every score of
    // the step (DESIGN.md "mistake_prob"); 0 for the perfect user
    double score = 0.0;
    if (row < 0) { row = best->idx; score = best->score + shift_coef * h_base[1]; }
    if (ct.enabled && score != score) row = -1;      // a NaN score never wins (k_pick_winner)
    const int rec_len = 8 + w_cap + d;
    if (row < 0) {
        for (int k = threadIdx.x; k < rec_len; k += blockDim.x) rec[k] = k == 0 ? -1.0 : (k == 1 ? -INFINITY : 0.0);
        if (ct.enabled && threadIdx.x == 0) {
            ct.rec_in[0] = -1.0;
            ct.rec_in[1] = -INFINITY;
            ct.sel[2 * ct.t] = -1.0;
            ct.sel[2 * ct.t + 1] = -INFINITY;
        }
        if (pp.peer_base != nullptr) peer_put(pp, rec, rec_len);
        return;
    }
    for (int j = threadIdx.x; j < w_cap; j += blockDim.x) rec[8 + j] = j < W_tot ? U[(int64_t)j * ldu + row] : 0.0;
    for (int j = threadIdx.x; j < d; j += blockDim.x) rec[8 + w_cap + j] = (double)X[row * (int64_t)d_pad + j];
    if (threadIdx.x == 0) {
        double cv = v[row];
        for (int j = W_lab; j < W_tot; ++j) { const double e = U[(int64_t)j * ldu + row]; cv = fma(-e, e, cv); }
        rec[0] = (double)(row_offset + row);
        rec[1] = score;
        rec[2] = m[row];
        rec[3] = cv;
        rec[4] = sqn[row];
        rec[5] = v[row];
        rec[6] = gain[row];
        rec[7] = 0.0;
    }
    if (ct.enabled) {       // single shard: this record is the step's winner -- commit it here (k_pick_winner, G = 1)
        __syncthreads();
        for (int k = threadIdx.x; k < rec_len; k += blockDim.x) {
            ct.rec_in[k] = rec[k];
            ct.rec_hist_t[k] = rec[k];
        }
        if (threadIdx.x == 0) {
            if (row < ct.n) ct.mask[row] |= ct.mark_bits;
            ct.base_m[ct.t] = rec[2];
            for (int j = 0; j < ct.t; ++j) ct.base_L[ct.t * kBaseStride + j] = rec[8 + W_lab + j];
            ct.base_L[ct.t * kBaseStride + ct.t] = sqrt(fmax(rec[3], 1e-300));
            ct.sel[2 * ct.t] = rec[0];
            ct.sel[2 * ct.t + 1] = rec[1];
        }
    }
    if (pp.peer_base != nullptr) peer_put(pp, rec, rec_len);
}

// ---- GaussianProcess.update without the host (ital/gp.py:164-200): the model lives on the device ----------------
// The Cholesky factor of K_LL + noise I (row-major, leading dimension ldk), beta = L_K^-1 y and the labelled rows
// (as float64) are appended on the device, so that update() needs no device-to-host round trip.
//
// k_prepare_labelled: one block.  For the q <= 4 new points (local rows idx[a], targets y[a]) it gathers what the
// labelled pass needs from the pool -- the rows as float64, their projections on the current factor, posterior mean
// and variance -- computes the q x q triangle of the block Cholesky extension among them
//   T[a][b] = (k(x_a, x_b) - u_a . u_b - sum_{c<b} T[a][c] T[b][c]) / piv_b,  piv_a^2 = v_a - sum_b T[a][b]^2 + noise,
// writes the MultiExt block (and, for q = 1, the point record k_extend reads), appends the model rows and marks the
// points as seen.  Dot products are reduced in a fixed order (thread-strided partial sums, xor butterfly, warps in
// ascending order).
__device__ __forceinline__ double block_sum_256(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    return t;
}

struct ModelRefs {                // device-resident model
    double* LK;                   // [cap][ldk] lower triangle
    int64_t ldk;
    double* beta;
    double* lab_x;                // [cap][d]
    double* lab_sqn;
};

// append rows W .. W+q-1 of the model from a MultiExt block (header, z[q][d_pad], ur[q][W]); every thread of one block
__device__ __forceinline__ void append_model_rows(const ModelRefs& M, int q, int W, int d, int d_pad, const double* ext) {
    const MultiExt* h = reinterpret_cast<const MultiExt*>(ext);
    const double* z = ext + sizeof(MultiExt) / sizeof(double);
    const double* ur = z + (size_t)q * d_pad;
    for (int a = 0; a < q; ++a) {
        double* row = M.LK + (int64_t)(W + a) * M.ldk;
        for (int j = threadIdx.x; j < W; j += blockDim.x) row[j] = ur[(size_t)a * W + j];
        for (int j = threadIdx.x; j < d; j += blockDim.x) M.lab_x[(int64_t)(W + a) * d + j] = z[(size_t)a * d_pad + j];
        if (threadIdx.x == 0) {
            for (int b = 0; b < a; ++b) row[W + b] = h->tri[a * 4 + b];
            row[W + a] = h->piv[a];
            M.beta[W + a] = h->beta[a];
            M.lab_sqn[W + a] = h->zn[a];
        }
    }
}

__global__ void __launch_bounds__(256) k_append_model(ModelRefs M, int q, int W, int d, int d_pad,
                                                      const double* __restrict__ ext) {
    pdl_enter();
    append_model_rows(M, q, W, d, d_pad, ext);
}

template <typename XT>
__global__ void __launch_bounds__(256) k_prepare_labelled(int q, const int64_t* __restrict__ idx,
                                                          const double* __restrict__ yv, int64_t row_offset, int64_t n,
                                                          const XT* __restrict__ X, int d, int d_pad,
                                                          const double* __restrict__ sqn, const double* __restrict__ m,
                                                          const double* __restrict__ v, const double* __restrict__ U,
                                                          int64_t ldu, int W, int w_cap, double var, double neg2ls2,
                                                          double noise, double* __restrict__ ext,
                                                          double* __restrict__ rec1, ModelRefs M,
                                                          uint8_t* __restrict__ mask, uint8_t seen_bits) {
    pdl_enter();
    __shared__ double red[8];
    __shared__ double hm[4], hv[4], tri[16], piv[4], beta[4];
    MultiExt* h = reinterpret_cast<MultiExt*>(ext);
    double* z = ext + sizeof(MultiExt) / sizeof(double);
    double* ur = z + (size_t)q * d_pad;
    for (int a = 0; a < q; ++a) {
        const int64_t row = idx[a] - row_offset;
        for (int j = threadIdx.x; j < d_pad; j += blockDim.x)
            z[(size_t)a * d_pad + j] = j < d ? (double)X[row * (int64_t)d_pad + j] : 0.0;
        for (int j = threadIdx.x; j < W; j += blockDim.x) ur[(size_t)a * W + j] = U[(int64_t)j * ldu + row];
        if (threadIdx.x == 0) {
            hm[a] = m[row];
            hv[a] = v[row];
            h->zn[a] = sqn[row];
            mask[row] |= seen_bits;
        }
    }
    if (threadIdx.x < 16) tri[threadIdx.x] = 0.0;
    __syncthreads();
    for (int a = 0; a < q; ++a) {
        double cv = hv[a], ma = hm[a];
        for (int b = 0; b < a; ++b) {
            double pd = 0.0, pu = 0.0;
            for (int j = threadIdx.x; j < d; j += blockDim.x) pd = fma(z[(size_t)a * d_pad + j], z[(size_t)b * d_pad + j], pd);
            for (int j = threadIdx.x; j < W; j += blockDim.x) pu = fma(ur[(size_t)a * W + j], ur[(size_t)b * W + j], pu);
            const double dot = block_sum_256(pd, red);
            const double proj = block_sum_256(pu, red);
            double num = var * exp((h->zn[a] + h->zn[b] - 2.0 * dot) / neg2ls2) - proj;
            for (int c = 0; c < b; ++c) num -= tri[a * 4 + c] * tri[b * 4 + c];
            const double e = num / piv[b];
            __syncthreads();
            if (threadIdx.x == 0) tri[a * 4 + b] = e;
            cv -= e * e;
            ma += e * beta[b];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            piv[a] = sqrt(fmax(cv + noise, 2.3e-308));
            beta[a] = (yv[a] - ma) / piv[a];
        }
        __syncthreads();
    }
    if (threadIdx.x < 16) h->tri[threadIdx.x] = tri[threadIdx.x];
    if (threadIdx.x < 4) {
        h->piv[threadIdx.x] = threadIdx.x < q ? piv[threadIdx.x] : 0.0;
        h->beta[threadIdx.x] = threadIdx.x < q ? beta[threadIdx.x] : 0.0;
        if ((int)threadIdx.x >= q) h->zn[threadIdx.x] = 0.0;
    }
    if (q == 1 && rec1 != nullptr) {                    // the single-column pass reads a point record (k_extend)
        for (int j = threadIdx.x; j < w_cap; j += blockDim.x) rec1[8 + j] = j < W ? ur[j] : 0.0;
        for (int j = threadIdx.x; j < d; j += blockDim.x) rec1[8 + w_cap + j] = z[j];
        if (threadIdx.x == 0) {
            rec1[0] = (double)idx[0];
            rec1[1] = 0.0;
            rec1[2] = hm[0];
            rec1[3] = hv[0];
            rec1[4] = h->zn[0];
            rec1[5] = hv[0];
            rec1[6] = 0.0;
            rec1[7] = 0.0;
        }
    }
    __syncthreads();
    append_model_rows(M, q, W, d, d_pad, ext);
}

// w = K^-1 y = L_K^-T beta (gp.py:158,196) by back substitution in one block: column sweep, the row of L_K that is
// eliminated is contiguous.
__global__ void __launch_bounds__(1024) k_model_w(ModelRefs M, int W, double* __restrict__ w) {
    pdl_enter();
    extern __shared__ double wsm[];
    for (int j = threadIdx.x; j < W; j += blockDim.x) wsm[j] = M.beta[j];
    __syncthreads();
    for (int a = W - 1; a >= 0; --a) {
        const double* row = M.LK + (int64_t)a * M.ldk;
        const double wa = wsm[a] / row[a];
        __syncthreads();
        for (int b = threadIdx.x; b < a; b += blockDim.x) wsm[b] = fma(-row[b], wa, wsm[b]);
        if (threadIdx.x == 0) wsm[a] = wa;
        __syncthreads();
    }
    for (int j = threadIdx.x; j < W; j += blockDim.x) w[j] = wsm[j];
}

// GaussianProcess.predict (ital/gp.py:264-292) for arbitrary rows: one warp per test row.
//   k_l = var * exp((|x|^2 + |x_l|^2 - 2 x . x_l) / s);  mean = w . k;  var = max(0, var - |L_K^-1 k|^2)
__global__ void __launch_bounds__(128) k_predict(const double* __restrict__ Xt, int64_t mrows, int d,
                                                 const double* __restrict__ Xl, const double* __restrict__ sqn_l,
                                                 int nl, const double* __restrict__ wvec,
                                                 const double* __restrict__ LK, int64_t ldk, double var, double neg2ls2,
                                                 double* __restrict__ out_mean, double* __restrict__ out_var,
                                                 double* __restrict__ out_proj) {
    pdl_enter();
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* kbuf = sm + (size_t)wib * nl;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (row >= mrows) return;
    const double* x = Xt + row * (int64_t)d;
    double xn = 0.0;
    for (int j = lane; j < d; j += 32) xn = fma(x[j], x[j], xn);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xn += __shfl_xor_sync(0xffffffffu, xn, o);
    for (int l = 0; l < nl; ++l) {
        double acc = 0.0;
        const double* xl = Xl + (int64_t)l * d;
        for (int j = lane; j < d; j += 32) acc = fma(x[j], xl[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) kbuf[l] = var * exp((sqn_l[l] + xn - 2.0 * acc) / neg2ls2);
    }
    __syncwarp();
    if (lane == 0) {
        double mean = 0.0;
        for (int l = 0; l < nl; ++l) mean = fma(wvec[l], kbuf[l], mean);
        out_mean[row] = mean;
        if (out_var != nullptr || out_proj != nullptr) {
            double q = 0.0;
            for (int a = 0; a < nl; ++a) {               // forward substitution with the Cholesky factor
                double u = kbuf[a];
                for (int b = 0; b < a; ++b) u = fma(-LK[(int64_t)a * ldk + b], kbuf[b], u);
                u /= LK[(int64_t)a * ldk + a];
                kbuf[a] = u;
                q = fma(u, u, q);
                if (out_proj != nullptr) out_proj[row * (int64_t)nl + a] = u;
            }
            if (out_var != nullptr) out_var[row] = fmax(0.0, var - q);
        }
    }
}

// ---- ActiveRetrievalBase.top_results (ital/retrieval_base.py:64-75): np.argsort(rel_mean)[::-1] on the device ----
// Stable LSD radix sort of (key, row) pairs, 8 bits per pass, keys = posterior means mapped to unsigned integers
// whose ascending order is the descending order of the means (exact ties keep ascending row order).  Per pass:
// k_sort_hist (digit histogram of every tile), k_sort_scan (per digit: exclusive scan over the tiles + digit total),
// k_sort_scatter (stable ranks inside a tile by warp match + per-warp counters).
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                                   // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;             // 4096 keys per block

__device__ __forceinline__ uint64_t desc_key(double x) {
    const uint64_t b = (uint64_t)__double_as_longlong(x);
    const uint64_t asc = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    return ~asc;
}

// `mask` != nullptr: rows with a non-zero mask (seen, not candidates) get the largest key, i.e. they sort behind every
// candidate (the top_candidates restriction ranks the unseen rows only, ital/ital.py:116)
__global__ void __launch_bounds__(256) k_sort_init(const double* __restrict__ m, int64_t n, uint64_t* __restrict__ keys,
                                                   uint32_t* __restrict__ rows, const uint8_t* __restrict__ mask) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = (mask != nullptr && mask[i] != 0) ? ~0ull : desc_key(m[i]);
        rows[i] = (uint32_t)i;
    }
}

// the first `top` rows of a sorted row list become the only candidates: clears `clear_bits` in their masks
__global__ void __launch_bounds__(256) k_mask_clear_sorted(uint8_t* __restrict__ mask, const uint32_t* __restrict__ rows,
                                                           int64_t top, uint8_t clear_bits) {
    pdl_enter();
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < top; k += (int64_t)gridDim.x * blockDim.x)
        mask[rows[k]] &= (uint8_t)~clear_bits;
}

// item (w, j, lane) of a tile: tile_base + w * (32 * kSortItems) + j * 32 + lane -- the order that defines stability
__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                            uint32_t* __restrict__ hist) {
    pdl_enter();
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int j = 0; j < kSortItems; ++j) {
        const int64_t i = base + (int64_t)j * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// One block per digit: exclusive scan of the digit's tile counts in place, digit total to totals[digit].
__global__ void __launch_bounds__(256) k_sort_scan(uint32_t* __restrict__ hist, int tiles, uint32_t* __restrict__ totals) {
    pdl_enter();
    __shared__ uint32_t wsum[8];
    uint32_t* h = hist + (size_t)blockIdx.x * tiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int base = 0; base < tiles; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t c = i < tiles ? h[i] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        uint32_t before = carry, all = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) {
            if (ww < w) before += wsum[ww];
            all += wsum[ww];
        }
        if (i < tiles) h[i] = before + inc - c;
        carry += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint64_t* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ rows_in, int64_t n,
                                                               int shift, const uint32_t* __restrict__ offs,
                                                               const uint32_t* __restrict__ totals,
                                                               uint64_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ rows_out) {
    pdl_enter();
    constexpr int kWarps = kSortThreads / 32;
    __shared__ uint32_t cnt[kWarps][256];
    __shared__ uint32_t dsum[kWarps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < kWarps * 256; k += kSortThreads) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile + (int64_t)w * 32 * kSortItems;
    uint64_t key[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const int64_t i = base + j * 32 + lane;
        const bool ok = i < n;
        key[j] = ok ? keys_in[i] : ~0ull;
        const uint32_t dgt = ok ? (uint32_t)((key[j] >> shift) & 255u) : 256u;       // 256: not an item
        const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && lane == leader) { old = cnt[w][dgt]; cnt[w][dgt] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // thread = digit: exclusive scan over the warps of the tile, on top of the tile's global offset
        // (keys with smaller digits anywhere + keys with this digit in earlier tiles)
        const uint32_t tot = totals[threadIdx.x];
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) dsum[w] = inc;
        __syncthreads();
        uint32_t run = inc - tot + offs[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
        for (int ww = 0; ww < w; ++ww) run += dsum[ww];
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) { const uint32_t c = cnt[ww][threadIdx.x]; cnt[ww][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const int64_t i = base + j * 32 + lane;
        if (i < n) {
            const uint32_t pos = cnt[w][(key[j] >> shift) & 255u] + rank[j];
            keys_out[pos] = key[j];
            rows_out[pos] = rows_in[i];
        }
    }
}

__global__ void __launch_bounds__(256) k_sort_gather(const uint32_t* __restrict__ rows, int64_t k, int64_t row_offset,
                                                     const double* __restrict__ m, int64_t* __restrict__ out_idx,
                                                     double* __restrict__ out_val) {
    pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = rows[i];
        out_idx[i] = row_offset + (int64_t)r;
        out_val[i] = m[r];
    }
}

}  // namespace italk
