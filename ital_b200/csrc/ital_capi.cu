// C ABI of the ITAL batch-selection path (include/ital_b200.h): shard state, launch sequencing, small
// host-side linear algebra (the |L| x |L| Cholesky factor grows by one row per labelled point and stays on the
// host; everything that scales with the pool runs in the kernels of ital_kernels.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/ital_b200.h"
#include "ital_kernels.cuh"
#include "ital_fused.cuh"
#include "snq_host.h"

using namespace italk;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ITAL_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr uint8_t kSeen = 1, kSelected = 2, kNotCandidate = 4, kRestricted = 8, kUnnameable = 16;
// Slack of the lazy-greedy bound: a row is pruned only if its bound (a gain evaluated exactly at an EARLIER step) stays
// below the best exact score of this step by more than this, so it has to cover the quadrature error of both steps'
// node sets (measured against converged rules, tests/test_orthant_vs_scipy.py: scores to 1e-7 for up to 4 base
// variables, 4e-6 for 5, 1e-4 for the lattice from 6 on).
constexpr double kPruneMargin = 1e-6;
inline double prune_margin(int t) { return t <= 3 ? kPruneMargin : (t == 4 ? 2e-6 : (t == 5 ? 2e-5 : 1e-3)); }
constexpr int kArgmaxBlocks = 592;      // 4 x 148 SMs

}  // namespace

struct ital_shard {
    int device = 0;
    cudaStream_t stream = nullptr;
    int x_dtype = ITAL_F32;
    int64_t n = 0, d = 0, d_pad = 0, row_offset = 0, n_data = 0, ldu = 0;
    double ls = 0, var = 1, noise = 1e-6;
    int num_sms = 148;

    void* X = nullptr;
    double *sqn = nullptr, *m = nullptr, *v = nullptr, *U = nullptr, *gain = nullptr, *score = nullptr;
    uint8_t* mask = nullptr;
    int* worklist = nullptr;
    int* counters = nullptr;         // [0] worklist size, [1] flagged, [2] scored exactly
    Best* block_best = nullptr;      // kArgmaxBlocks
    Best* best = nullptr;            // [0] step winner, [1] best of stage A
    double* thr_dev = nullptr;       // worklist threshold (device scalar)
    double* rec_dev = nullptr;       // records produced on the device (propose / export)
    double* rec_host = nullptr;      // pinned mirror of rec_dev
    double* rec_in_dev = nullptr;    // the record k_extend reads (one record)
    double* rec_in_host = nullptr;   // pinned staging of rec_in_dev
    int64_t rec_cap = 0;
    int64_t* idx_dev = nullptr;      // scratch for index lists
    int64_t idx_cap = 0;
    // quadrature nodes of the current step
    double *eta_dev = nullptr, *w_dev = nullptr, *masses_dev = nullptr;
    int *group_dev = nullptr, *orth_dev = nullptr;
    double *eta_raw = nullptr, *w_raw = nullptr;     // nodes as generated, before the negligible ones are dropped
    int* orth_raw = nullptr;
    int64_t nodes_cap = 0;
    int64_t n_nodes = 0;
    double* gl_dev = nullptr;        // Gauss-Legendre tables: x[65][64] then w[65][64]
    double2* phi_dev = nullptr;      // (Phi, phi) on the grid of phi_tab
    double* htab_dev = nullptr;      // polynomial table of h_tab (closed-form score of the first greedy step)
    // batch state on the device: means and Cholesky rows of the selected points, selection list, H(base)
    double *base_m_dev = nullptr, *base_L_dev = nullptr, *sel_dev = nullptr, *hbase_dev = nullptr;
    double* sel_host = nullptr;      // pinned mirror of sel_dev: (global row, score) per step
    int* stats_dev = nullptr;        // per step: worklist size, flagged, scored, -
    int* stats_host = nullptr;       // pinned
    int proposals = 0;               // propose calls in the running fetch
    bool lazy_rows = false;          // batch projections only for the rows that get scored (k_catchup)
    // peer-memory exchange of the per-step proposals (multi-GPU fetch without NCCL in the loop)
    int xg_world = 0, xg_rank = 0;
    int64_t xg_slot = 0;                     // doubles per slot
    unsigned char* xg_local = nullptr;       // this shard's exchange buffer (exported through CUDA IPC)
    std::vector<unsigned char*> xg_peer;     // mapped buffers of all shards ([rank] = xg_local)
    unsigned char** xg_peer_dev = nullptr;
    int* xg_error_dev = nullptr;
    unsigned long long xg_epoch = 0;
    bool xg_ready = false;
    // work of the NEXT greedy step that does not depend on the streaming pass (quadrature nodes, stage-A bound
    // argmax) runs on a side stream while the pass runs on all SMs but one (ITAL_B200_OVERLAP=0: off)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_commit = nullptr, ev_side = nullptr;
    int nodes_ready_t = -1, stage_a_ready_t = -1;
    cudaEvent_t ev_win[16] = {};             // pipelined fetch: winner of step t committed
    const double* ext_rec = nullptr;         // record the streaming pass reads (default: rec_in_dev)
    bool ahead_cols = false;                 // scoring runs beside the pass that writes the newest column
    int reserve_sms = 4;                     // SMs the pass leaves at least to the side stream (tools: ITAL_B200_RESERVE)
    bool overlap = true, last_exhaustive = false, reserve_sm = false;
    PickSrc pick;                    // where k_record finds the local best of the running step
    bool pdl = true;                 // programmatic dependent launch between the kernels of a stream (ITAL_B200_PDL=0: off)
    bool bulk_stream = true;         // X stream staged by the bulk-copy engine (k_extend_bulk) where it applies
    uint32_t* tags = nullptr;        // per row: [fetch epoch | batch columns valid | step of the last exact score]
    uint32_t epoch = 0;              // epoch of the running fetch (tags of older epochs read as empty)
    // the fused persistent fetch kernel (k_fetch_fused): scratch, grid barrier, switches
    Best* f_best = nullptr;
    int* f_cnt = nullptr;
    int* f_stage = nullptr;
    double* f_mass = nullptr;
    unsigned* f_bar = nullptr;
    unsigned f_bar_count = 0;        // arrivals on the barrier counter so far (monotone across launches)
    bool fused = true;               // ITAL_B200_FUSED=0: multi-kernel loop only
    unsigned long long* f_trace = nullptr;   // phase time stamps of the last fused launch (ital_fused_trace)
    bool f_trace_on = false;
    bool sel_marked = false;         // rows of the running batch carry the kSelected mask bit
    bool hbase_seeded = false;       // hbase_dev holds {0, 1} for the first greedy step of the running fetch
    int fused_steps_last = 0;        // greedy steps the fused kernel ran in the last fetch (diagnostics)
    double* rec_hist = nullptr;      // records of the points selected in the running fetch
    uint64_t* sort_keys = nullptr;   // top_results: two key and two row buffers (ping-pong), tile histograms
    uint32_t* sort_rows = nullptr;
    uint32_t* sort_hist = nullptr;
    int64_t* sort_out_idx = nullptr;
    double* sort_out_val = nullptr;
    int64_t sort_cap = 0;
    double* mext_dev = nullptr;      // block of the multi-column labelled extension (MultiExt + z + ur)
    double* mext_host = nullptr;     // pinned
    size_t mext_cap = 0;
    int64_t rec_hist_cap = 0;
    // general feedback model (label_prob < 1): conditional node sets of the running step
    double *g_eta = nullptr, *g_w = nullptr, *g_mass = nullptr;
    int *g_begin = nullptr, *g_set0 = nullptr, *g_lut = nullptr;
    size_t g_cap_nodes = 0, g_cap_groups = 0, g_cap_sets = 0, g_cap_lut = 0;
    // change_estimation_subset: node sets of ital_fetch_propose_sub
    double *sb_eta = nullptr, *sb_w = nullptr, *sb_small = nullptr;
    int* sb_begin = nullptr;
    size_t sb_cap_eta = 0, sb_cap_w = 0, sb_cap_small = 0, sb_cap_begin = 0;
    // sequential-conditioning lattice generated on the device (t >= 6): scratch
    double* sc_dbl = nullptr;        // P[1024] | chunk_sum[4096]
    int* sc_int = nullptr;           // cnt[1024] | chunk_orth[4096] | chunk_off[4096] | n_chunks
    int* tn_int = nullptr;           // tensor rule on the device (t = 4, 5): kept nodes per block and orthant, their scan
    size_t tn_cap = 0;
    bool device_lattice = true;      // (ITAL_B200_DEVICE_LATTICE=0: host generation, for A/B comparisons)
    // clip_cov (grouped orthant probabilities from the sixth sample of a batch on)
    double clip_cov = 0.0;
    unsigned short* clip_sub = nullptr;  // [n] union of the batch's groups a candidate is connected to
    int* clip_int = nullptr;             // present[32] | count2 | comp_mask[16] | gb (grown)
    double* clip_dbl = nullptr;          // sd_b[16] | mass | T (grown)
    int* clip_list2 = nullptr;           // [n]
    void* clip_desc = nullptr;           // ClipDesc[1024]
    double *clip_eta = nullptr, *clip_w = nullptr;
    size_t clip_cap_int = 0, clip_cap_dbl = 0, clip_cap_eta = 0, clip_cap_w = 0;
    bool sub_mode = false;           // the batch columns hold ext = [batch, subset]: no look-ahead for a next greedy step

    int w_cap = 0;                   // allocated projection columns
    int W = 0;                       // labelled points in the model
    int t = 0;                       // points selected in the running fetch
    bool fetching = false;
    double label_prob = 1.0, mistake_prob = 0.0;
    int estimation = 0;              // label_estimation: 0 'mean', 1 'optimistic', 2 'pessimistic' (ital/ital.py:210-219)

    // the small model lives on the device (appended by k_prepare_labelled / k_append_model): Cholesky factor of
    // K_LL + noise I (row-major lower triangle, leading dimension model_cap), beta = L_K^-1 y, the labelled rows as
    // float64 and their squared norms; w = K^-1 y is derived on demand (k_model_w) for predict()
    double *lab_x_dev = nullptr, *lab_sqn_dev = nullptr, *w_vec_dev = nullptr, *LK_dev = nullptr, *beta_dev = nullptr;
    int model_cap = 0;
    bool w_valid = false;
    int64_t* upd_idx_dev = nullptr;          // parameters of ital_update_labelled: rows and targets of the new points
    double* upd_y_dev = nullptr;

    double stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double step_nodes[16] = {0};
    // measurement
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    double prof_bytes = 0.0;
    int64_t launches = 0;
    double log1p_eps = std::log(1.0 + 1e-12);
    int64_t h2d_bytes = 0, d2h_bytes = 0;      // host <-> device bytes copied by the library since creation
};

namespace {

// every host <-> device copy of the library goes through these two, so that the bytes a call moves can be read back
// (ital_transfer_bytes; bench.py reports them per fetch instead of a hand-kept constant)
cudaError_t copy_async(ital_shard* s, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    if (kind == cudaMemcpyHostToDevice) s->h2d_bytes += (int64_t)bytes;
    else if (kind == cudaMemcpyDeviceToHost) s->d2h_bytes += (int64_t)bytes;
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
}
cudaError_t copy_sync(ital_shard* s, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    if (kind == cudaMemcpyHostToDevice) s->h2d_bytes += (int64_t)bytes;
    else if (kind == cudaMemcpyDeviceToHost) s->d2h_bytes += (int64_t)bytes;
    return cudaMemcpy(dst, src, bytes, kind);
}

}  // namespace

namespace {

int64_t record_doubles(const ital_shard* s) { return ITAL_RECORD_HEADER + s->w_cap + s->d; }

// H(u) = -Phi(u) log(Phi(u) + eps) - Phi(-u) log(Phi(-u) + eps) as degree-7 polynomials per interval of width 1/16
// over [0, 8.5] (h_tab in ital_kernels.cuh): interpolation at the Chebyshev nodes of every interval, in extended
// precision; coefficients in ascending powers of the position x in [-1, 1] inside the interval.
// (Phi, phi) on the grid x_k = -8.5 + k/128 of phi_tab, in extended precision
std::vector<double> build_phi_table() {
    std::vector<double> tab((size_t)kPhiTableLen * 2);
    for (int k = 0; k < kPhiTableLen; ++k) {
        const long double x = -(long double)kPhiXMax + (long double)k / kPhiPerUnit;
        tab[(size_t)k * 2 + 0] = (double)(0.5L * erfcl(-x * 0.70710678118654752440084436210484903L));
        tab[(size_t)k * 2 + 1] = (double)(expl(-0.5L * x * x) * 0.39894228040143267793994605993438187L);
    }
    return tab;
}

std::vector<double> build_h_table() {
    std::vector<double> tab((size_t)kHTabLen * 8);
    const long double eps = 1e-12L, pi = 3.14159265358979323846264338327950288L;
    auto H = [&](long double u) {
        const long double p1 = 0.5L * erfcl(-u * 0.70710678118654752440084436210484903L);
        const long double p0 = 0.5L * erfcl(u * 0.70710678118654752440084436210484903L);
        return -(p1 * logl(p1 + eps) + p0 * logl(p0 + eps));
    };
    for (int k = 0; k < kHTabLen; ++k) {
        const long double mid = ((long double)k + 0.5L) / kHTabPerUnit, half = 0.5L / kHTabPerUnit;
        long double A[8][9];
        for (int i = 0; i < 8; ++i) {
            const long double x = cosl(pi * (2 * i + 1) / 16.0L);
            long double pw = 1.0L;
            for (int j = 0; j < 8; ++j) { A[i][j] = pw; pw *= x; }
            A[i][8] = H(mid + half * x);
        }
        for (int c = 0; c < 8; ++c) {       // Gauss-Jordan with partial pivoting
            int piv = c;
            for (int r = c + 1; r < 8; ++r)
                if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
            for (int j = 0; j < 9; ++j) std::swap(A[c][j], A[piv][j]);
            for (int r = 0; r < 8; ++r) {
                if (r == c) continue;
                const long double f = A[r][c] / A[c][c];
                for (int j = c; j < 9; ++j) A[r][j] -= f * A[c][j];
            }
        }
        for (int j = 0; j < 8; ++j) tab[(size_t)k * 8 + j] = (double)(A[j][8] / A[j][j]);
    }
    return tab;
}


// Kernel launch with programmatic stream serialization (see pdl_enter in ital_kernels.cuh): the kernel may start
// launching while its predecessor in the stream drains; it waits for the predecessor's completion itself.  All
// arguments are given explicitly (default arguments do not exist behind a function pointer).
template <typename... KArgs>
struct PdlLaunch {
    void (*kernel)(KArgs...);
    dim3 grid, block;
    size_t smem;
    cudaStream_t stream;
    bool programmatic;
    template <typename... Args>
    void operator()(Args&&... args) const {
        static_assert(sizeof...(Args) == sizeof...(KArgs), "pass every kernel argument explicitly");
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = programmatic ? 1 : 0;
        cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);   // errors: cudaGetLastError
    }
};

template <typename... KArgs>
PdlLaunch<KArgs...> pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, const ital_shard* s) {
    return PdlLaunch<KArgs...>{kernel, grid, block, smem, s->stream, s->pdl};
}

int grid_for(const ital_shard* s, int64_t work_items, int per_block, int max_waves = 8) {
    int64_t blocks = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)s->num_sms * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int ensure_record_buffers(ital_shard* s, int q) {
    const int64_t need = record_doubles(s) * q;
    if (need <= s->rec_cap) return ITAL_OK;
    CU(cudaStreamSynchronize(s->stream));
    if (s->rec_dev) CU(cudaFree(s->rec_dev));
    if (s->rec_host) CU(cudaFreeHost(s->rec_host));
    if (s->rec_in_dev) CU(cudaFree(s->rec_in_dev));
    if (s->rec_in_host) CU(cudaFreeHost(s->rec_in_host));
    s->rec_dev = s->rec_host = s->rec_in_dev = s->rec_in_host = nullptr;
    CU(cudaMalloc(&s->rec_dev, need * sizeof(double)));
    CU(cudaMallocHost(&s->rec_host, need * sizeof(double)));
    CU(cudaMalloc(&s->rec_in_dev, record_doubles(s) * sizeof(double)));
    CU(cudaMallocHost(&s->rec_in_host, record_doubles(s) * sizeof(double)));
    s->rec_cap = need;
    return ITAL_OK;
}

int ensure_idx(ital_shard* s, int64_t m) {
    if (m <= s->idx_cap) return ITAL_OK;
    if (s->idx_dev) CU(cudaFree(s->idx_dev));
    s->idx_dev = nullptr;
    CU(cudaMalloc(&s->idx_dev, m * sizeof(int64_t)));
    s->idx_cap = m;
    return ITAL_OK;
}

// grow U to hold at least `cols` projection columns
int ensure_width(ital_shard* s, int cols) {
    if (cols <= s->w_cap) return ITAL_OK;
    // 64 columns to start with: a session of up to ~50 labelled points never reallocates (growing U means a new
    // allocation of n x cap doubles and a device-to-device copy, 3 to 600 ms at n = 10^6 depending on the allocator)
    int new_cap = std::max(64, s->w_cap);
    while (new_cap < cols) new_cap *= 2;
    double* nu = nullptr;
    CU(cudaMalloc(&nu, (size_t)new_cap * s->ldu * sizeof(double)));
    if (s->U) {
        CU(copy_async(s, nu, s->U, (size_t)s->w_cap * s->ldu * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        CU(cudaFree(s->U));
    }
    s->U = nu;
    s->w_cap = new_cap;
    // records change size with the capacity
    s->rec_cap = 0;
    return ensure_record_buffers(s, 1);
}

// CTAs of a persistent streaming kernel whose warps own whole 32-row units in a static round-robin: the count in
// [lo, hi] that wastes the least of the last round (with 10^6 rows and 16 warps per CTA, 148 CTAs leave 80 % of the
// warps idle in round 14 of 14, while 140 CTAs fill 13.95 of 14 rounds; the pass is HBM-bound, not SM-bound, so
// fewer, evenly loaded CTAs are faster).  Ties go to the larger count.
int balanced_ctas(int64_t units, int warps_per_cta, int lo, int hi) {
    lo = std::max(1, lo);
    hi = std::max(lo, hi);
    int best = hi;
    double best_eff = -1.0;
    for (int c = hi; c >= lo; --c) {
        const int64_t w = (int64_t)c * warps_per_cta;
        const int64_t rounds = (units + w - 1) / w;
        const double eff = (double)units / (double)(rounds * w);
        if (eff > best_eff + 1e-12) { best_eff = eff; best = c; }
    }
    return best;
}

template <typename XT>
int launch_extend_t(ital_shard* s, int W_used, int labelled, double y, uint8_t mark_bits) {
    constexpr int VN = Vec<XT>::N;
    const int nchunks = (int)(s->d_pad / (32 * VN));
    const int threads = 256, warps = threads / 32;
    const bool fixed = (nchunks == 1 || nchunks == 2 || nchunks == 4);
    size_t smem = ((size_t)warps * 32 * 33 + ((W_used + 1) & ~1) + (fixed ? 0 : s->d_pad)) * sizeof(double);
    const int64_t units = (s->n + 31) / 32;
    int blocks = (int)std::min<int64_t>((units + warps - 1) / warps, (int64_t)s->num_sms * ITAL_EXTEND_MINB);
    if (blocks < 1) blocks = 1;
    const double neg2ls2 = -2.0 * (s->ls * s->ls);
    const double* ext_rec = s->ext_rec ? s->ext_rec : s->rec_in_dev;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (s->profiling) {
        CU(cudaEventCreate(&ev0));
        CU(cudaEventCreate(&ev1));
        CU(cudaEventRecord(ev0, s->stream));
    }
    if (s->bulk_stream && nchunks == 4 && s->d_pad * sizeof(XT) == 2048) {    // 2 KB rows: the tuned shape
        // TMA-staged variant: one CTA per SM, every warp with a private ring of slots
        const int bthreads = kBulkThreads, bwarps = bthreads / 32;
        const size_t ring = (size_t)bwarps * kBulkSlots * kBulkRows * s->d_pad * sizeof(XT);
        const size_t bsmem = ring + (size_t)((W_used + 1) & ~1) * sizeof(double) +
                             (size_t)bwarps * kBulkSlots * sizeof(uint64_t);
        const int hi = s->num_sms - (s->reserve_sm ? s->reserve_sms : 0);
        const int bblocks = (units + bwarps - 1) / bwarps <= hi ? (int)std::max<int64_t>(1, (units + bwarps - 1) / bwarps)
                                                                : balanced_ctas(units, bwarps, hi - 12, hi);
#define ITAL_LAUNCH_BULK(NCV)                                                                                     \
    do {                                                                                                          \
        CU(cudaFuncSetAttribute(k_extend_bulk<XT, NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem)); \
        pdl(k_extend_bulk<XT, NCV>, bblocks, bthreads, bsmem, s)(                                           \
            (const XT*)s->X, s->n, (int)s->d, (int)s->d_pad, ext_rec, s->w_cap, W_used, s->sqn, s->U,       \
            s->ldu, s->m, s->v, labelled, y, s->noise, s->var, neg2ls2);                                          \
    } while (0)
        if (nchunks == 4) ITAL_LAUNCH_BULK(4);
        else if (nchunks == 2) ITAL_LAUNCH_BULK(2);
        else ITAL_LAUNCH_BULK(1);
#undef ITAL_LAUNCH_BULK
    } else {
#define ITAL_LAUNCH_EXT(NCV)                                                                                   \
    do {                                                                                                       \
        CU(cudaFuncSetAttribute(k_extend<XT, NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        pdl(k_extend<XT, NCV>, blocks, threads, smem, s)(                                                \
            (const XT*)s->X, s->n, (int)s->d, (int)s->d_pad, ext_rec, s->w_cap, W_used, s->sqn, s->U,     \
            s->ldu, s->m, s->v, labelled, y, s->noise, s->var, neg2ls2, s->mask, s->row_offset, mark_bits);     \
    } while (0)
    if (nchunks == 4) ITAL_LAUNCH_EXT(4);
    else if (nchunks == 2) ITAL_LAUNCH_EXT(2);
    else if (nchunks == 1) ITAL_LAUNCH_EXT(1);
    else ITAL_LAUNCH_EXT(0);
#undef ITAL_LAUNCH_EXT
    }
    s->launches++;
    CU(cudaGetLastError());
    if (s->profiling) {
        CU(cudaEventRecord(ev1, s->stream));
        s->prof_events.emplace_back(ev0, ev1);
        // algorithmic bytes of one pass: the row, |x|^2, W projections in, one out; m and v
        // read-modify-write when a labelled point is added
        s->prof_bytes += (double)s->n * ((double)s->d * sizeof(XT) + 8.0 + 8.0 * W_used + 8.0 + (labelled ? 32.0 : 0.0));
    }
    return ITAL_OK;
}

// One streaming pass: extend every local row's projection by the point described by the host record `rec`
// (current layout).  Writes column `col`, using the first `col` entries of the record's projection.  The record
// goes through pinned staging, so nothing here waits for the device.
int extend_with_record(ital_shard* s, const double* rec, int col, int labelled, double y, bool mark_selected) {
    int rc = ensure_width(s, col + 1);
    if (rc) return rc;
    const int64_t rl = record_doubles(s);
    memcpy(s->rec_in_host, rec, (size_t)rl * sizeof(double));
    CU(copy_async(s, s->rec_in_dev, s->rec_in_host, (size_t)rl * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    const uint8_t mark = mark_selected ? kSelected : (uint8_t)0;
    if (s->x_dtype == ITAL_F32) return launch_extend_t<float>(s, col, labelled, y, mark);
    return launch_extend_t<double>(s, col, labelled, y, mark);
}


// room for `rows` labelled points in the device-resident model (grows by doubling; the old rows are copied over)
int ensure_model(ital_shard* s, int rows) {
    if (rows <= s->model_cap) return ITAL_OK;
    int cap = std::max(64, s->model_cap);
    while (cap < rows) cap *= 2;
    double *nx = nullptr, *nq = nullptr, *nw = nullptr, *nl = nullptr, *nb = nullptr;
    CU(cudaMalloc(&nx, (size_t)cap * s->d * sizeof(double)));
    CU(cudaMalloc(&nq, (size_t)cap * sizeof(double)));
    CU(cudaMalloc(&nw, (size_t)cap * sizeof(double)));
    CU(cudaMalloc(&nl, (size_t)cap * cap * sizeof(double)));
    CU(cudaMalloc(&nb, (size_t)cap * sizeof(double)));
    if (s->W > 0 && s->model_cap > 0) {
        CU(cudaMemcpyAsync(nx, s->lab_x_dev, (size_t)s->W * s->d * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaMemcpyAsync(nq, s->lab_sqn_dev, (size_t)s->W * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaMemcpyAsync(nb, s->beta_dev, (size_t)s->W * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        CU(cudaMemcpy2DAsync(nl, (size_t)cap * sizeof(double), s->LK_dev, (size_t)s->model_cap * sizeof(double),
                             (size_t)s->W * sizeof(double), (size_t)s->W, cudaMemcpyDeviceToDevice, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    for (double* p : {s->lab_x_dev, s->lab_sqn_dev, s->w_vec_dev, s->LK_dev, s->beta_dev})
        if (p) CU(cudaFree(p));
    s->lab_x_dev = nx;
    s->lab_sqn_dev = nq;
    s->w_vec_dev = nw;
    s->LK_dev = nl;
    s->beta_dev = nb;
    s->model_cap = cap;
    s->w_valid = false;
    return ITAL_OK;
}

ModelRefs model_refs(const ital_shard* s) {
    ModelRefs M;
    M.LK = s->LK_dev;
    M.ldk = s->model_cap;
    M.beta = s->beta_dev;
    M.lab_x = s->lab_x_dev;
    M.lab_sqn = s->lab_sqn_dev;
    return M;
}

// device + pinned block of the multi-column labelled extension: MultiExt header, z[q][d_pad], ur[q][W]
int ensure_mext(ital_shard* s, int q, int W) {
    const size_t need = sizeof(MultiExt) / sizeof(double) + (size_t)q * s->d_pad + (size_t)q * W;
    if (need <= s->mext_cap) return ITAL_OK;
    CU(cudaStreamSynchronize(s->stream));
    if (s->mext_dev) CU(cudaFree(s->mext_dev));
    if (s->mext_host) CU(cudaFreeHost(s->mext_host));
    s->mext_dev = s->mext_host = nullptr;
    const size_t cap = need * 2;
    CU(cudaMalloc(&s->mext_dev, cap * sizeof(double)));
    CU(cudaMallocHost(&s->mext_host, cap * sizeof(double)));
    s->mext_cap = cap;
    return ITAL_OK;
}

template <typename XT>
int launch_extend_multi(ital_shard* s, int q, int W_used) {
    const int threads = 256, warps = threads / 32;
    const size_t smem = ((size_t)q * s->d_pad + (size_t)q * W_used) * sizeof(double);
    const int64_t units = (s->n + 31) / 32;
    int blocks = (int)std::min<int64_t>((units + warps - 1) / warps, (int64_t)s->num_sms * 2);
    if (blocks < 1) blocks = 1;
    const double neg2ls2 = -2.0 * (s->ls * s->ls);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (s->profiling) {
        CU(cudaEventCreate(&ev0));
        CU(cudaEventCreate(&ev1));
        CU(cudaEventRecord(ev0, s->stream));
    }
    constexpr int VN = Vec<XT>::N;
    const int bwarps = kMultiThreads / 32;
    const size_t bsmem = (size_t)bwarps * kMultiSlots * kMultiRows * s->d_pad * sizeof(XT) +
                         ((size_t)q * s->d_pad + (size_t)((q * W_used + 1) & ~1)) * sizeof(double) +
                         (size_t)bwarps * kMultiSlots * sizeof(uint64_t);
    if (s->bulk_stream && s->d_pad * sizeof(XT) == 2048 && s->d_pad / (32 * VN) == 4 && bsmem <= 227 * 1024) {
        // 2 KB rows: the TMA-staged variant, one CTA per SM
        const int bblocks = (units + bwarps - 1) / bwarps <= s->num_sms ? (int)std::max<int64_t>(1, (units + bwarps - 1) / bwarps)
                                                                        : balanced_ctas(units, bwarps, s->num_sms - 12, s->num_sms);
#define ITAL_LAUNCH_BMULTI(QV)                                                                                        \
    do {                                                                                                              \
        CU(cudaFuncSetAttribute(k_extend_bulk_multi<XT, 4, QV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem)); \
        pdl(k_extend_bulk_multi<XT, 4, QV>, bblocks, kMultiThreads, bsmem, s)(                                  \
            (const XT*)s->X, s->n, (int)s->d_pad, s->mext_dev, W_used, s->sqn, s->U, s->ldu, s->m, s->v, s->var,      \
            neg2ls2);                                                                                                 \
    } while (0)
        if (q == 2) ITAL_LAUNCH_BMULTI(2);
        else if (q == 3) ITAL_LAUNCH_BMULTI(3);
        else ITAL_LAUNCH_BMULTI(4);
#undef ITAL_LAUNCH_BMULTI
    } else {
#define ITAL_LAUNCH_MULTI(QV)                                                                                     \
    do {                                                                                                          \
        CU(cudaFuncSetAttribute(k_extend_multi<XT, QV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pdl(k_extend_multi<XT, QV>, blocks, threads, smem, s)((const XT*)s->X, s->n, (int)s->d_pad,       \
                                                                      s->mext_dev, W_used, s->sqn, s->U, s->ldu,  \
                                                                      s->m, s->v, s->var, neg2ls2);               \
    } while (0)
    if (q == 2) ITAL_LAUNCH_MULTI(2);
    else if (q == 3) ITAL_LAUNCH_MULTI(3);
    else ITAL_LAUNCH_MULTI(4);
#undef ITAL_LAUNCH_MULTI
    }
    s->launches++;
    CU(cudaGetLastError());
    if (s->profiling) {
        CU(cudaEventRecord(ev1, s->stream));
        s->prof_events.emplace_back(ev0, ev1);
        s->prof_bytes += (double)s->n * ((double)s->d * sizeof(XT) + 8.0 + 8.0 * W_used + 8.0 * q + 32.0);
    }
    return ITAL_OK;
}

// additive constant of the scores of the running step for a user who mislabels with probability mistake_prob and
// always answers (label_prob >= 1): every relevance configuration r keeps log(1 + eps) with probability
// (1 - mp)^(t+1) and gets log(eps) otherwise (the updated orthant probability is 0 after a contradicting label), so
// score = perfect-user score + (1 - (1 - mp)^(t+1)) * (log eps - log(1 + eps)) * sum_r p_r.
double step_shift_coef(const ital_shard* s) {
    if (!(s->mistake_prob > 0.0) || s->label_prob < 1.0 || s->estimation != 0) return 0.0;
    const double c = std::pow(1.0 - s->mistake_prob, (double)(s->t + 1));
    return (1.0 - c) * (std::log(1e-12) - s->log1p_eps);
}

int make_record(ital_shard* s, long long local_row, double* dst_dev, bool commit_here = false, PickSrc src = PickSrc(),
                PeerPut pp = PeerPut()) {
    const double shift = (local_row < 0 && !s->sub_mode) ? step_shift_coef(s) : 0.0;     // (k_eval_sub scores in full)
    CommitTargets ct;
    if (commit_here) {
        s->sel_marked = true;
        ct.rec_in = s->rec_in_dev;
        ct.rec_hist_t = s->rec_hist + (int64_t)s->t * record_doubles(s);
        ct.base_m = s->base_m_dev;
        ct.base_L = s->base_L_dev;
        ct.sel = s->sel_dev;
        ct.mask = s->mask;
        ct.n = s->n;
        ct.t = s->t;
        ct.mark_bits = kSelected;
        ct.enabled = 1;
    }
    if (s->x_dtype == ITAL_F32)
        pdl(k_record<float>, 1, 256, 0, s)(local_row, s->best, s->row_offset, (const float*)s->X, (int)s->d,
                                                  (int)s->d_pad, s->sqn, s->m, s->v, s->U, s->ldu, s->W,
                                                  s->W + s->t, s->w_cap, s->gain, dst_dev, shift, s->hbase_dev,
                                                  s->counters, local_row < 0 ? s->stats_dev + 4 * s->t : nullptr, ct, src, pp);
    else
        pdl(k_record<double>, 1, 256, 0, s)(local_row, s->best, s->row_offset, (const double*)s->X,
                                                   (int)s->d, (int)s->d_pad, s->sqn, s->m, s->v, s->U, s->ldu,
                                                   s->W, s->W + s->t, s->w_cap, s->gain, dst_dev, shift, s->hbase_dev,
                                                   s->counters, local_row < 0 ? s->stats_dev + 4 * s->t : nullptr, ct, src, pp); s->launches++;
    CU(cudaGetLastError());
    return ITAL_OK;
}

constexpr int kMaxBatch = 11;            // greedy steps per fetch (t <= 10 base variables)
constexpr int kMaxBatchGeneral = 5;      // ... with label_prob < 1 (conditional node sets up to 4 base variables)

int ensure_nodes(ital_shard* s, int64_t n_nodes) {
    n_nodes += 8 * kNodePad;                            // (every orthant of the packed node list is zero-padded to 256)
    if (n_nodes <= s->nodes_cap) return ITAL_OK;
    CU(cudaStreamSynchronize(s->stream));
    for (void* p : {(void*)s->eta_dev, (void*)s->w_dev, (void*)s->orth_dev, (void*)s->eta_raw, (void*)s->w_raw,
                    (void*)s->orth_raw})
        if (p) CU(cudaFree(p));
    s->eta_dev = s->w_dev = s->eta_raw = s->w_raw = nullptr;
    s->orth_dev = s->orth_raw = nullptr;
    CU(cudaMalloc(&s->eta_raw, (size_t)n_nodes * 3 * sizeof(double)));
    CU(cudaMalloc(&s->w_raw, (size_t)n_nodes * sizeof(double)));
    CU(cudaMalloc(&s->orth_raw, (size_t)n_nodes * sizeof(int)));
    CU(cudaMalloc(&s->eta_dev, (size_t)n_nodes * 10 * sizeof(double)));
    CU(cudaMalloc(&s->w_dev, (size_t)n_nodes * sizeof(double)));
    CU(cudaMalloc(&s->orth_dev, (size_t)n_nodes * sizeof(int)));
    s->nodes_cap = n_nodes;
    return ITAL_OK;
}

// Quadrature nodes of the step from the batch state.  t <= 3: generated on the device from the device-resident
// state, nothing waits.  t >= 4 (batches of more than 4): generated and sorted on the host (csrc/snq_host.h),
// which costs one device-to-host round trip for the batch state.
int prepare_nodes(ital_shard* s) {
    const int t = s->t;
    const int q = snq::order_for(t);
    const int64_t N = snq::capacity_for(t);
    int rc = ensure_nodes(s, N);
    if (rc) return rc;
    s->n_nodes = N;
    if (t <= 3) {
        const int blocks = (int)((N + 255) / 256);
        const double* glx = s->gl_dev;
        const double* glw = s->gl_dev + (snq::kMaxOrder + 1) * 64;
#define ITAL_GEN(TV) pdl(k_snq_generate<TV>, blocks, 256, 0, s)(q, snq::kR, snq::kQMin, s->base_m_dev, s->base_L_dev, \
                                                              glx, glw, N, s->eta_raw, s->w_raw, s->orth_raw)
        if (t == 1) ITAL_GEN(1);
        else if (t == 2) ITAL_GEN(2);
        else ITAL_GEN(3);
#undef ITAL_GEN
        s->launches++;
        const size_t fsm = ((size_t)s->num_sms * 8 + 8) * sizeof(double);
        pdl(k_snq_finalize, 1, 1024, fsm, s)(t, N, snq::kWMin, s->eta_raw, s->w_raw, s->orth_raw, s->eta_dev,
                                             s->group_dev, s->log1p_eps, s->masses_dev, s->hbase_dev, s->counters + 3,
                                             s->num_sms); s->launches++;
        CU(cudaGetLastError());
        return ITAL_OK;
    }
    if (t < snq::kScFrom && s->device_lattice) {
        // 4 or 5 base variables: the tensor rule, counted and scattered on the device
        const int nb = 1 << t;
        const int n_blocks = (int)((N + 255) / 256);
        if ((size_t)n_blocks * nb > s->tn_cap) {
            CU(cudaStreamSynchronize(s->stream));
            if (s->tn_int) CU(cudaFree(s->tn_int));
            s->tn_int = nullptr;
            CU(cudaMalloc(&s->tn_int, (size_t)n_blocks * nb * 2 * sizeof(int)));
            s->tn_cap = (size_t)n_blocks * nb;
        }
        int* blk_cnt = s->tn_int;
        int* blk_off = s->tn_int + (size_t)n_blocks * nb;
        const double* glx = s->gl_dev;
        const double* glw = s->gl_dev + (snq::kMaxOrder + 1) * 64;
        if (t == 4) {
            pdl(k_tn_count<4>, n_blocks, 256, 0, s)(q, snq::kR, snq::kQMin, s->base_m_dev, s->base_L_dev, glx, glw, N,
                                                    snq::kWMin, blk_cnt);
        } else {
            pdl(k_tn_count<5>, n_blocks, 256, 0, s)(q, snq::kR, snq::kQMin, s->base_m_dev, s->base_L_dev, glx, glw, N,
                                                    snq::kWMin, blk_cnt);
        }
        s->launches++;
        pdl(k_tn_scan, 1, 1024, 0, s)(nb, n_blocks, blk_cnt, blk_off, s->group_dev, s->counters + 3); s->launches++;
        if (t == 4) {
            pdl(k_tn_scatter<4>, n_blocks, 256, 0, s)(q, snq::kR, snq::kQMin, s->base_m_dev, s->base_L_dev, glx, glw, N,
                                                      snq::kWMin, blk_off, s->group_dev, N, s->eta_dev, s->w_dev);
        } else {
            pdl(k_tn_scatter<5>, n_blocks, 256, 0, s)(q, snq::kR, snq::kQMin, s->base_m_dev, s->base_L_dev, glx, glw, N,
                                                      snq::kWMin, blk_off, s->group_dev, N, s->eta_dev, s->w_dev);
        }
        s->launches++;
        pdl(k_tn_masses, nb, 256, 0, s)(s->group_dev, s->w_dev, s->masses_dev); s->launches++;
        pdl(k_tn_hbase, 1, 32, 0, s)(nb, s->masses_dev, s->log1p_eps, s->hbase_dev); s->launches++;
        CU(cudaGetLastError());
        return ITAL_OK;
    }
    if (t >= snq::kScFrom && s->device_lattice) {
        // the lattice inside every base orthant, generated where the batch state lives: nothing waits for the host
        if (!s->sc_dbl) {
            CU(cudaMalloc(&s->sc_dbl, (1024 + 4096) * sizeof(double)));
            CU(cudaMalloc(&s->sc_int, (1024 + 2 * 4096 + 1) * sizeof(int)));
        }
        static const int primes[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
        ScArgs a;
        a.t = t;
        for (int j = 0; j < 12; ++j) {
            const double r = std::sqrt((double)primes[j]);
            a.alpha[j] = r - std::floor(r);
        }
        a.base_m = s->base_m_dev;
        a.base_L = s->base_L_dev;
        a.n_total = snq::kScN;
        a.pilot = snq::kScPilot;
        a.n_min = snq::kScMin;
        a.p_min = snq::kScPMin;
        a.stride = N;
        a.P = s->sc_dbl;
        a.chunk_sum = s->sc_dbl + 1024;
        a.cnt = s->sc_int;
        a.chunk_orth = s->sc_int + 1024;
        a.chunk_off = s->sc_int + 1024 + 4096;
        a.n_chunks = s->sc_int + 1024 + 2 * 4096;
        a.group_begin = s->group_dev;
        a.eta = s->eta_dev;
        a.w = s->w_dev;
        a.masses = s->masses_dev;
        a.hbase = s->hbase_dev;
        a.log1p_eps = s->log1p_eps;
        const int nb = 1 << t;
        pdl(k_sc_pilot, nb, 256, 0, s)(a); s->launches++;
        pdl(k_sc_alloc, 1, 1024, 0, s)(a); s->launches++;
        pdl(k_sc_generate, std::min(4096, 8 * s->num_sms), 256, 0, s)(a); s->launches++;
        pdl(k_sc_masses, 1, 1024, 0, s)(a); s->launches++;
        pdl(k_sc_scale, 2 * s->num_sms, 256, 0, s)(a); s->launches++;
        CU(cudaGetLastError());
        return ITAL_OK;
    }
    // host path
    std::vector<double> bm(16), bL(16 * 16);
    CU(copy_async(s, bm.data(), s->base_m_dev, 16 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(copy_async(s, bL.data(), s->base_L_dev, 256 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    std::vector<double> Lb((size_t)t * t, 0.0);
    for (int a = 0; a < t; ++a)
        for (int b = 0; b <= a; ++b) Lb[(size_t)a * t + b] = bL[(size_t)a * kBaseStride + b];
    snq::Nodes nd = snq::generate(t, bm.data(), Lb.data());
    const int nb = 1 << t;
    s->n_nodes = nd.n;                      // stride of the dimension-major coordinates
    CU(copy_async(s, s->eta_dev, nd.eta.data(), nd.eta.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->w_dev, nd.w.data(), nd.w.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->masses_dev, nd.masses.data(), nb * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->group_dev, nd.group_begin.data(), (nb + 1) * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    double hb[2] = {nd.entropy, 0.0};
    for (double mval : nd.masses) hb[1] += mval;
    CU(copy_async(s, s->hbase_dev, hb, sizeof hb, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));   // nd lives in pageable host memory
    return ITAL_OK;
}

// lazy rows: make the listed rows' batch projections current before they are scored
// `ahead`: streaming mode, side stream -- the listed rows get the batch columns they still lack from the stored
// winner records, exactly as in lazy mode, while the streaming passes write the same columns for every row on the
// main stream (same bits, so whichever store lands last is immaterial)
int launch_catchup(ital_shard* s, int64_t items_hint, bool ahead = false) {
    if ((!s->lazy_rows && !ahead) || s->t == 0) return ITAL_OK;
    const int threads = 256, warps = threads / 32;
    const int blocks = grid_for(s, std::max<int64_t>(1, items_hint), warps, 8);
    const size_t smem = (size_t)warps * s->w_cap * sizeof(double);
    const double neg2ls2 = -2.0 * (s->ls * s->ls);
    if (s->x_dtype == ITAL_F32) {
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_catchup<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pdl(k_catchup<float>, blocks, threads, smem, s)(s->counters, s->worklist, (const float*)s->X, (int)s->d,
                                                        (int)s->d_pad, s->rec_hist, record_doubles(s), s->w_cap,
                                                        s->W, s->t, s->sqn, s->U, s->ldu, s->tags, s->epoch, s->var, neg2ls2);
    } else {
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_catchup<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pdl(k_catchup<double>, blocks, threads, smem, s)(s->counters, s->worklist, (const double*)s->X, (int)s->d,
                                                         (int)s->d_pad, s->rec_hist, record_doubles(s), s->w_cap,
                                                         s->W, s->t, s->sqn, s->U, s->ldu, s->tags, s->epoch, s->var, neg2ls2);
    }
    s->launches++;
    CU(cudaGetLastError());
    return ITAL_OK;
}

// Score the rows in the worklist.  `items_hint` bounds the number of items (the true count is on the device);
// few items -> one 256-thread block per candidate (latency), many -> one warp per candidate (throughput).
int launch_eval(ital_shard* s, int64_t items_hint, bool block_per_candidate) {
    int rc_c = launch_catchup(s, items_hint);
    if (rc_c) return rc_c;
    const int threads = 256;
    const int blocks = grid_for(s, std::max<int64_t>(1, items_hint), block_per_candidate ? 1 : threads / 32, 8);
    EvalArgs a;
    a.count = s->counters;
    a.list = s->worklist;
    a.m = s->m;
    a.v = s->v;
    a.U = s->U;
    a.ldu = s->ldu;
    a.W0 = s->W;
    a.eta = s->eta_dev;
    a.w = s->w_dev;
    a.nodes4 = s->eta_dev;
    a.orth = s->orth_dev;
    a.phi = s->phi_dev;
    a.group_begin = s->group_dev;
    a.n_nodes = s->n_nodes;
    a.n_kept = s->counters + 3;
    a.masses = s->masses_dev;
    a.h_base = s->hbase_dev;
    a.log1p_eps = s->log1p_eps;
    a.flag_var = 100.0 * s->noise;
    a.score = s->score;
    a.gain = s->gain;
    a.tags = s->tags;
    a.epoch = s->epoch;
    a.n_flagged = s->counters + 1;
    a.n_scored = s->counters + 2;
    a.force_block = block_per_candidate ? 1 : 0;
    a.t = s->t;
    if (s->t == 1) pdl(k_eval<1>, blocks, threads, 0, s)(a);
    else if (s->t == 2) pdl(k_eval<2>, blocks, threads, 0, s)(a);
    else if (s->t == 3) pdl(k_eval<3>, blocks, threads, 0, s)(a);
    else {
        pdl(k_eval_sorted, blocks, threads, 0, s)(a);
        pdl(k_eval_sorted_pair, blocks, threads, 0, s)(a); s->launches++;      // (one of the two returns at once)
    }
    s->launches++;
    CU(cudaGetLastError());
    return ITAL_OK;
}

// General feedback model, steps t >= 1: every candidate is scored (the lazy-greedy bound is only proven for users
// who label every sample); the conditional node sets come from the host, which costs one round trip per step.
int propose_general(ital_shard* s) {
    const int t = s->t;
    if (t >= kMaxBatchGeneral) return fail(ITAL_EINVAL, "label_prob < 1 supports batches of at most %d samples", kMaxBatchGeneral);
    std::vector<double> bm(16), bL(16 * 16);
    CU(copy_async(s, bm.data(), s->base_m_dev, 16 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(copy_async(s, bL.data(), s->base_L_dev, 256 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    std::vector<double> Lb((size_t)t * t, 0.0);
    for (int a = 0; a < t; ++a)
        for (int b = 0; b <= a; ++b) Lb[(size_t)a * t + b] = bL[(size_t)a * kBaseStride + b];
    snq::GeneralSets gs = snq::generate_general(t, bm.data(), Lb.data(), s->noise);
    auto grow = [&](void** p, size_t* cap, size_t need, size_t esize) -> int {
        if (need <= *cap) return ITAL_OK;
        if (*p) CU(cudaFree(*p));
        *p = nullptr;
        CU(cudaMalloc(p, need * esize));
        *cap = need;
        return ITAL_OK;
    };
    int rc;
    size_t cap_w = s->g_cap_nodes;
    if ((rc = grow((void**)&s->g_eta, &s->g_cap_nodes, (size_t)gs.n_nodes * 4, sizeof(double)))) return rc;
    if ((rc = grow((void**)&s->g_w, &cap_w, (size_t)gs.n_nodes * 4, sizeof(double)))) return rc;
    size_t cap_b = s->g_cap_groups;
    if ((rc = grow((void**)&s->g_mass, &s->g_cap_groups, (size_t)gs.n_groups + 1, sizeof(double)))) return rc;
    if ((rc = grow((void**)&s->g_begin, &cap_b, (size_t)gs.n_groups + 1, sizeof(int)))) return rc;
    if ((rc = grow((void**)&s->g_set0, &s->g_cap_sets, (size_t)gs.n_sets + 1, sizeof(int)))) return rc;
    if ((rc = grow((void**)&s->g_lut, &s->g_cap_lut, gs.lut.size(), sizeof(int)))) return rc;
    CU(copy_async(s, s->g_eta, gs.eta.data(), gs.eta.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->g_w, gs.w.data(), gs.w.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->g_mass, gs.group_mass.data(), gs.group_mass.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->g_begin, gs.group_begin.data(), gs.group_begin.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->g_set0, gs.set_group0.data(), gs.set_group0.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->g_lut, gs.lut.data(), gs.lut.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));       // gs lives in pageable host memory
    s->n_nodes = gs.n_nodes;
    pdl(k_worklist, grid_for(s, s->n, 256), 256, 0, s)(s->n, s->mask, s->gain, s->thr_dev, 1, s->counters, s->worklist); s->launches++;
    GeneralArgs a;
    a.count = s->counters;
    a.list = s->worklist;
    a.m = s->m;
    a.v = s->v;
    a.U = s->U;
    a.ldu = s->ldu;
    a.W0 = s->W;
    a.t = t;
    a.eta = s->g_eta;
    a.w = s->g_w;
    a.n_nodes = gs.n_nodes;
    a.group_begin = s->g_begin;
    a.group_mass = s->g_mass;
    a.set_group0 = s->g_set0;
    a.lut = s->g_lut;
    a.n_groups = gs.n_groups;
    a.n_sets = gs.n_sets;
    a.lp = s->label_prob;
    a.mp = s->mistake_prob;
    a.noise = s->noise;
    a.phi = s->phi_dev;
    a.score = s->score;
    a.gain = s->gain;
    a.tags = s->tags;
    a.epoch = s->epoch;
    a.n_scored = s->counters + 2;
    a.estimation = s->estimation;
    a.fb_kind = (s->label_prob >= 1.0 && s->mistake_prob <= 0.0) ? 0 : (s->label_prob >= 1.0 ? 1 : 2);
    if ((rc = launch_catchup(s, s->n))) return rc;
    const size_t smem = ((((size_t)3 * gs.n_groups + 2 * gs.n_sets + 1) & ~(size_t)1) + (size_t)gs.n_groups * 24 + 8) * sizeof(double) +
                        (size_t)kPhiTableLen * sizeof(double2) + (s->estimation != 0 ? (32 + 1024 + 7776) * sizeof(double) : 0);      // logp, logq, the 2^5 * 3^5 terms
    CU(cudaFuncSetAttribute(k_eval_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = grid_for(s, s->n, 1, 8);
    pdl(k_eval_general, blocks, 256, smem, s)(a); s->launches++;
    const int lb = std::min(kArgmaxBlocks, grid_for(s, s->n, 256));
    pdl(k_argmax_list, lb, 256, 0, s)(s->counters, s->worklist, s->score, s->block_best, nullptr, 0.0, 0.0, nullptr, nullptr); s->launches++;
    s->pick = PickSrc();
    s->pick.block_best = s->block_best;
    s->pick.nblocks = lb;
    CU(cudaGetLastError());
    return ITAL_OK;
}

// clip_cov, steps with more than 5 samples (t >= 5 base variables), users who label everything: every candidate is
// scored (the lazy-greedy bound is not proven for the clipped model), see k_clip_adj / k_eval_clip.
int propose_clip(ital_shard* s) {
    const int t = s->t;
    if (t > 10) return fail(ITAL_EINVAL, "clip_cov: batches of more than 11 samples are not supported");
    std::vector<double> bm(16), bL(16 * 16);
    CU(copy_async(s, bm.data(), s->base_m_dev, 16 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(copy_async(s, bL.data(), s->base_L_dev, 256 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    // covariance of the batch, its groups (connected components of |corr| > clip_cov)
    std::vector<double> C((size_t)t * t, 0.0), sd(16, 1.0);
    for (int a = 0; a < t; ++a)
        for (int b = 0; b <= a; ++b) {
            double acc = 0.0;
            for (int j = 0; j <= b; ++j) acc += bL[a * kBaseStride + j] * bL[b * kBaseStride + j];
            C[a * t + b] = C[b * t + a] = acc;
        }
    for (int a = 0; a < t; ++a) sd[a] = std::sqrt(std::max(C[a * t + a], 1e-300));
    std::vector<int> comp(t);
    for (int a = 0; a < t; ++a) comp[a] = a;
    auto find = [&](int a) { while (comp[a] != a) a = comp[a] = comp[comp[a]]; return a; };
    for (int a = 0; a < t; ++a)
        for (int b = 0; b < a; ++b)
            if (std::fabs(C[a * t + b] / (sd[a] * sd[b])) > s->clip_cov) comp[find(a)] = find(b);
    std::vector<int> comp_mask(16, 0);
    for (int a = 0; a < t; ++a)
        for (int b = 0; b < t; ++b)
            if (find(a) == find(b)) comp_mask[a] |= 1 << b;
    // node set of a subset of the batch (members in batch order): marginal means and Cholesky factor
    auto subset_nodes = [&](int mask, std::vector<int>& mem, std::vector<double>& Ls) {
        mem.clear();
        for (int a = 0; a < t; ++a)
            if ((mask >> a) & 1) mem.push_back(a);
        const int u = (int)mem.size();
        std::vector<double> ms(u);
        Ls.assign((size_t)u * u, 0.0);
        for (int a = 0; a < u; ++a) {
            ms[a] = bm[mem[a]];
            for (int b = 0; b < u; ++b) Ls[a * u + b] = C[mem[a] * t + mem[b]];
        }
        snq::chol_inplace(Ls, u);
        return snq::generate(u, ms.data(), Ls.data());
    };
    // entropy of every group
    double h_all = 0.0;
    std::vector<double> h_comp(t, 0.0);
    std::vector<int> mem;
    std::vector<double> Ls;
    for (int a = 0; a < t; ++a) {
        const int lowest = __builtin_ctz(comp_mask[a]);         // (one evaluation per group: at its lowest member)
        if (lowest != a) continue;
        snq::Nodes nd = subset_nodes(comp_mask[a], mem, Ls);
        h_comp[a] = nd.entropy;
        h_all += nd.entropy;
    }
    // pass 1: the groups every candidate is connected to
    auto grow = [&](void** p, size_t* cap, size_t need, size_t esize) -> int {
        if (need <= *cap) return ITAL_OK;
        CU(cudaStreamSynchronize(s->stream));
        if (*p) CU(cudaFree(*p));
        *p = nullptr;
        CU(cudaMalloc(p, need * esize));
        *cap = need;
        return ITAL_OK;
    };
    if (!s->clip_sub) {
        CU(cudaMalloc(&s->clip_sub, (size_t)s->n * sizeof(unsigned short)));
        CU(cudaMalloc(&s->clip_list2, (size_t)s->n * sizeof(int)));
        CU(cudaMalloc(&s->clip_desc, 1024 * sizeof(ClipDesc)));
    }
    int rc;
    if ((rc = grow((void**)&s->clip_int, &s->clip_cap_int, 64, sizeof(int)))) return rc;
    if ((rc = grow((void**)&s->clip_dbl, &s->clip_cap_dbl, 64, sizeof(double)))) return rc;
    std::vector<int> head(64, 0);
    for (int a = 0; a < 16; ++a) head[33 + a] = comp_mask[a];
    CU(copy_async(s, s->clip_int, head.data(), 64 * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->clip_dbl, sd.data(), 16 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemsetAsync(s->counters, 0, 4 * sizeof(int), s->stream));
    pdl(k_worklist, grid_for(s, s->n, 256), 256, 0, s)(s->n, s->mask, s->gain, s->thr_dev, 1, s->counters, s->worklist); s->launches++;
    int rcu = launch_catchup(s, s->n);
    if (rcu) return rcu;
    ClipArgs a = {};
    a.count = s->counters;
    a.list = s->worklist;
    a.count2 = s->clip_int + 32;
    a.list2 = s->clip_list2;
    a.m = s->m;
    a.v = s->v;
    a.U = s->U;
    a.ldu = s->ldu;
    a.W0 = s->W;
    a.t = t;
    a.base_L = s->base_L_dev;
    a.sd_b = s->clip_dbl;
    a.comp_mask = s->clip_int + 33;
    a.th = s->clip_cov;
    a.h_all = h_all;
    a.sub = s->clip_sub;
    a.present = (unsigned*)s->clip_int;
    a.phi = s->phi_dev;
    a.log1p_eps = s->log1p_eps;
    a.score = s->score;
    a.gain = s->gain;
    a.tags = s->tags;
    a.epoch = s->epoch;
    a.n_scored = s->counters + 2;
    pdl(k_clip_adj, grid_for(s, s->n, 256), 256, 0, s)(a); s->launches++;
    std::vector<unsigned> present(33);
    CU(copy_async(s, present.data(), s->clip_int, 33 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    // pass 2: node sets of the unions that occur
    std::vector<ClipDesc> desc(1024);
    std::vector<double> eta, w, dbl(16, 0.0);
    std::vector<int> ints(64, 0);
    for (int a2 = 0; a2 < 16; ++a2) dbl[a2] = sd[a2];
    int64_t total_nodes = 0;
    for (int S = 1; S < (1 << t); ++S) {
        if (!((present[S >> 5] >> (S & 31)) & 1u)) continue;
        snq::Nodes nd = subset_nodes(S, mem, Ls);
        const int u = (int)mem.size();
        total_nodes += nd.n;
        if (total_nodes > (int64_t)48 << 20)
            return fail(ITAL_EINVAL, "clip_cov: the candidates connect %d and more different unions of groups (over 48M quadrature nodes)", S);
        ClipDesc& d = desc[S];
        d.eta_off = (long long)eta.size();
        d.n = (int)nd.n;
        d.dims = u;
        d.gb_off = (int)ints.size() - 64;
        d.mass_off = (int)dbl.size() - 16;
        ints.insert(ints.end(), nd.group_begin.begin(), nd.group_begin.end());
        dbl.insert(dbl.end(), nd.masses.begin(), nd.masses.end());
        d.T_off = (int)dbl.size() - 16;
        // T = Ls^-1 L_b[S, :]  (u x t): a candidate's projection on the set's own factor from its batch columns
        for (int r = 0; r < u; ++r)
            for (int c = 0; c < t; ++c) dbl.push_back(0.0);
        double* T = dbl.data() + 16 + d.T_off;
        for (int c = 0; c < t; ++c)
            for (int r = 0; r < u; ++r) {
                double val = c <= mem[r] ? bL[mem[r] * kBaseStride + c] : 0.0;
                for (int k = 0; k < r; ++k) val -= Ls[r * u + k] * T[k * t + c];
                T[r * t + c] = val / Ls[r * u + r];
            }
        d.h_rest = 0.0;
        for (int b = 0; b < t; ++b)
            if (__builtin_ctz(comp_mask[b]) == b && !(S & comp_mask[b])) d.h_rest += h_comp[b];
        // coordinates of the set (dimension-major, stride = its own node count) and its weights, each in one array
        // shared by all sets: offsets eta_off / pad
        eta.insert(eta.end(), nd.eta.begin(), nd.eta.end());
        d.pad = (int)w.size();
        w.insert(w.end(), nd.w.begin(), nd.w.end());
    }
    if ((rc = grow((void**)&s->clip_eta, &s->clip_cap_eta, eta.size() + 1, sizeof(double)))) return rc;
    if ((rc = grow((void**)&s->clip_w, &s->clip_cap_w, w.size() + 1, sizeof(double)))) return rc;
    // (the small tables are re-uploaded whole: head of 64 ints / 16 doubles, then the per-set pieces)
    for (int k2 = 0; k2 < 64; ++k2) ints[k2] = k2 < 33 ? (int)present[k2] : head[k2];
    if ((rc = grow((void**)&s->clip_int, &s->clip_cap_int, ints.size(), sizeof(int)))) return rc;
    if ((rc = grow((void**)&s->clip_dbl, &s->clip_cap_dbl, dbl.size(), sizeof(double)))) return rc;
    CU(copy_async(s, s->clip_int, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->clip_dbl, dbl.data(), dbl.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (!eta.empty()) CU(copy_async(s, s->clip_eta, eta.data(), eta.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (!w.empty()) CU(copy_async(s, s->clip_w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->clip_desc, desc.data(), desc.size() * sizeof(ClipDesc), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    a.count2 = s->clip_int + 32;
    a.comp_mask = s->clip_int + 33;
    a.present = (unsigned*)s->clip_int;
    a.sd_b = s->clip_dbl;
    a.desc = (const ClipDesc*)s->clip_desc;
    a.eta = s->clip_eta;
    a.w = s->clip_w;
    a.gb = s->clip_int + 64;
    a.mass = s->clip_dbl + 16;
    a.T = s->clip_dbl + 16;
    pdl(k_eval_clip, grid_for(s, s->n, 1, 8), 256, 0, s)(a); s->launches++;
    {   // {H(batch), total mass}: the mistaken user's additive constant is per unit of total mass (make_record)
        const double hb2[2] = {h_all, 1.0};
        memcpy(s->sel_host + 30, hb2, sizeof hb2);      // (pinned scratch at the tail of the selection mirror)
        CU(copy_async(s, s->hbase_dev, s->sel_host + 30, sizeof hb2, cudaMemcpyHostToDevice, s->stream));
    }
    const int lb = std::min(kArgmaxBlocks, grid_for(s, s->n, 256));
    pdl(k_argmax_list, lb, 256, 0, s)(s->counters, s->worklist, s->score, s->block_best, nullptr, 0.0, 0.0, nullptr, nullptr); s->launches++;
    s->pick = PickSrc();
    s->pick.block_best = s->block_best;
    s->pick.nblocks = lb;
    s->n_nodes = total_nodes;
    CU(cudaGetLastError());
    return ITAL_OK;
}

// change_estimation_subset: score the local candidates against ext = the D batch columns of the running fetch, the
// first tB of which are the samples picked so far (k_eval_sub, snq::generate_sub).  only_local >= 0: that row alone.
int propose_sub(ital_shard* s, int tB, int64_t only_local, bool any_local) {
    const int D = s->t;
    std::vector<double> bm(16), bL(16 * 16);
    CU(copy_async(s, bm.data(), s->base_m_dev, 16 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(copy_async(s, bL.data(), s->base_L_dev, 256 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    std::vector<double> Lb((size_t)D * D, 0.0);
    for (int a = 0; a < D; ++a)
        for (int b = 0; b <= a; ++b) Lb[(size_t)a * D + b] = bL[(size_t)a * kBaseStride + b];
    snq::SubSets ss = snq::generate_sub(tB, D, bm.data(), Lb.data(), s->noise);
    const int G = 1 << tB;
    auto grow = [&](void** p, size_t* cap, size_t need, size_t esize) -> int {
        if (need <= *cap) return ITAL_OK;
        CU(cudaStreamSynchronize(s->stream));
        if (*p) CU(cudaFree(*p));
        *p = nullptr;
        CU(cudaMalloc(p, need * esize));
        *cap = need;
        return ITAL_OK;
    };
    int rc;
    // small tables in one buffer: mass [2G] | mu [G D] | Sig [D D] | mU [G u] | CU [u u] | BS [u D]
    const int u = D - tB;
    if (u > kSubMaxU) return fail(ITAL_EINVAL, "change_estimation_subset: at most %d subset members outside the batch", kSubMaxU);
    std::vector<double> small_tab;
    small_tab.insert(small_tab.end(), ss.mass.begin(), ss.mass.end());
    small_tab.insert(small_tab.end(), ss.mu.begin(), ss.mu.end());
    small_tab.insert(small_tab.end(), ss.Sig.begin(), ss.Sig.end());
    const size_t off_mU = small_tab.size();
    small_tab.insert(small_tab.end(), ss.mU.begin(), ss.mU.end());
    const size_t off_CU = small_tab.size();
    small_tab.insert(small_tab.end(), ss.CU.begin(), ss.CU.end());
    const size_t off_BS = small_tab.size();
    small_tab.insert(small_tab.end(), ss.BS.begin(), ss.BS.end());
    small_tab.push_back(0.0);
    if ((rc = grow((void**)&s->sb_eta, &s->sb_cap_eta, ss.eta.size() + 1, sizeof(double)))) return rc;
    if ((rc = grow((void**)&s->sb_w, &s->sb_cap_w, ss.w.size() + 1, sizeof(double)))) return rc;
    if ((rc = grow((void**)&s->sb_small, &s->sb_cap_small, small_tab.size(), sizeof(double)))) return rc;
    if ((rc = grow((void**)&s->sb_begin, &s->sb_cap_begin, ss.group_begin.size(), sizeof(int)))) return rc;
    CU(copy_async(s, s->sb_eta, ss.eta.data(), ss.eta.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->sb_w, ss.w.data(), ss.w.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->sb_small, small_tab.data(), small_tab.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->sb_begin, ss.group_begin.data(), ss.group_begin.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));       // ss lives in pageable host memory
    s->n_nodes = ss.n_nodes;
    CU(cudaMemsetAsync(s->counters, 0, 4 * sizeof(int), s->stream));
    if (only_local >= 0 || !any_local) {
        // a single row (if this shard owns it and it is a candidate), or nothing at all on this shard
        const int one[2] = {(int)only_local, 0};
        memcpy(s->sel_host + 30, one, sizeof one);      // (pinned scratch at the tail of the selection mirror)
        if (only_local >= 0) {
            CU(copy_async(s, s->worklist, s->sel_host + 30, sizeof(int), cudaMemcpyHostToDevice, s->stream));
            pdl(k_worklist_one, 1, 32, 0, s)(s->mask, s->worklist, s->counters); s->launches++;
        }
    } else {
        pdl(k_worklist, grid_for(s, s->n, 256), 256, 0, s)(s->n, s->mask, s->gain, s->thr_dev, 1, s->counters, s->worklist); s->launches++;
    }
    SubArgs a;
    a.count = s->counters;
    a.list = s->worklist;
    a.m = s->m;
    a.v = s->v;
    a.U = s->U;
    a.ldu = s->ldu;
    a.W0 = s->W;
    a.tB = tB;
    a.D = D;
    a.eta = s->sb_eta;
    a.w = s->sb_w;
    a.n_nodes = ss.n_nodes;
    a.group_begin = s->sb_begin;
    a.mass = s->sb_small;
    a.mu = s->sb_small + 2 * G;
    a.Sig = s->sb_small + 2 * G + (size_t)G * D;
    a.mU = s->sb_small + off_mU;
    a.CU = s->sb_small + off_CU;
    a.BS = s->sb_small + off_BS;
    a.sub_bits = ss.sub_bits;
    a.c1 = std::pow(1.0 - s->mistake_prob, (double)(tB + 1));
    a.q_last = u >= 2 ? snq::order_for(u - 1) : 1;
    a.R = snq::kR;
    a.q_min = snq::kQMin;
    a.gl_x = s->gl_dev;
    a.gl_w = s->gl_dev + (snq::kMaxOrder + 1) * 64;
    a.noise = s->noise;
    a.phi = s->phi_dev;
    a.score = s->score;
    a.gain = s->gain;
    a.tags = s->tags;
    a.epoch = s->epoch;
    a.n_scored = s->counters + 2;
    const size_t smem = (size_t)((4 * G + 16 + 5 * kBaseStride + 1) & ~1) * sizeof(double) + (size_t)kPhiTableLen * sizeof(double2);
    CU(cudaFuncSetAttribute(k_eval_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = only_local >= 0 ? 1 : grid_for(s, s->n, 1, 8);
    pdl(k_eval_sub, blocks, 256, smem, s)(a); s->launches++;
    const int lb = std::min(kArgmaxBlocks, grid_for(s, s->n, 256));
    pdl(k_argmax_list, lb, 256, 0, s)(s->counters, s->worklist, s->score, s->block_best, nullptr, 0.0, 0.0, nullptr, nullptr); s->launches++;
    s->pick = PickSrc();
    s->pick.block_best = s->block_best;
    s->pick.nblocks = lb;
    CU(cudaGetLastError());
    s->step_nodes[std::min(s->t, 15)] = (double)ss.n_nodes;
    return make_record(s, -1, s->rec_dev, false, s->pick);
}

// Stage A of a greedy step (t >= 1, perfect user, pruned): the bound argmax over 2 x #SM strided subsets, the exact
// scores of those rows, and from the best of them the threshold and the worklist of stage B.  None of it needs the
// streaming pass of the previous step except the new projection column of the ~300 stage-A rows themselves, which
// `ahead` computes on demand (k_catchup), so the whole stage can run beside the pass on the side stream.
int stage_a(ital_shard* s, double floor_score, bool ahead) {
    const int blocks = std::min(kArgmaxBlocks, grid_for(s, s->n, 256));
    const int ba = std::min(kArgmaxBlocks, std::min(blocks, 2 * s->num_sms));
    pdl(k_argmax_rows, ba, 512, 0, s)(s->n, s->gain, s->mask, s->block_best, s->counters + 4, s->counters,
                                      s->worklist); s->launches++;
    CU(cudaGetLastError());
    int rc = ahead ? launch_catchup(s, ba, true) : ITAL_OK;
    if (rc) return rc;
    rc = launch_eval(s, ba, true);
    if (rc) return rc;
    // best of stage A and, in the same kernel, the threshold of stage B: every row whose bound still
    // reaches the best exact score found so far
    pdl(k_argmax_list, 1, 256, 0, s)(s->counters, s->worklist, s->score, s->best + 1, s->hbase_dev,
                                     floor_score, prune_margin(s->t), s->thr_dev, s->counters); s->launches++;
    pdl(k_worklist, grid_for(s, s->n, 256), 256, 0, s)(s->n, s->mask, s->gain, s->thr_dev, 0,
                                                       s->counters, s->worklist); s->launches++;
    CU(cudaGetLastError());
    return ITAL_OK;
}

// The local candidates of the current greedy step -> record of the local best in DEVICE memory `rec_out`.
// Nothing here waits for the GPU (t <= 3).
int propose_dev(ital_shard* s, double floor_score, int exhaustive, double* rec_out, bool commit_here = false,
                PeerPut pp = PeerPut()) {
    if (s->t >= kMaxBatch) return fail(ITAL_EINVAL, "batches of more than %d samples are not supported", kMaxBatch);
    const int blocks = std::min(kArgmaxBlocks, grid_for(s, s->n, 256));
    s->pick = PickSrc();
    // counters [0..2] are zero here: k_record re-arms them at the end of every step, ital_fetch_begin before the first
    if (s->t == 0) CU(cudaMemsetAsync(s->counters, 0, 4 * sizeof(int), s->stream));
    if (s->t == 0 && !s->hbase_seeded) {
        const double hb0[2] = {0.0, 1.0};               // first step: no base, total mass 1
        memcpy(s->sel_host + 30, hb0, sizeof hb0);      // (pinned scratch at the tail of the selection mirror)
        CU(copy_async(s, s->hbase_dev, s->sel_host + 30, sizeof hb0, cudaMemcpyHostToDevice, s->stream));
        s->hbase_seeded = true;
    }
    if (s->t == 0 && s->estimation == 0) {
        const bool general = s->label_prob < 1.0;
        const double lc = general ? (1.0 - s->mistake_prob) * s->log1p_eps + s->mistake_prob * std::log(1e-12) : s->log1p_eps;
        pdl(k_score0, blocks, 256, 0, s)(s->n, s->m, s->v, s->mask, s->score, s->gain, s->block_best, lc,
                                                general ? s->label_prob : 1.0, s->htab_dev); s->launches++;
        s->pick.block_best = s->block_best;     // reduced by k_record
        s->pick.nblocks = blocks;
        CU(cudaGetLastError());
        s->n_nodes = 1;
    } else if (s->label_prob < 1.0 || s->estimation != 0) {
        int rc = propose_general(s);
        if (rc) return rc;
    } else if (s->clip_cov > 0.0 && s->clip_cov < 1.0 && s->t + 1 > 5) {
        int rc = propose_clip(s);                       // ital.py:360: grouped probabilities for more than 5 samples
        if (rc) return rc;
    } else {
        int rc = ITAL_OK;
        if (s->nodes_ready_t == s->t) s->n_nodes = snq::capacity_for(s->t);     // generated during the last pass
        else rc = prepare_nodes(s);
        if (rc) return rc;
        if (exhaustive) {
            pdl(k_worklist, grid_for(s, s->n, 256), 256, 0, s)(s->n, s->mask, s->gain, s->thr_dev, 1,
                                                                        s->counters, s->worklist); s->launches++;
            CU(cudaGetLastError());
            rc = launch_eval(s, s->n, false);
            if (rc) return rc;
            const int lb = std::min(kArgmaxBlocks, grid_for(s, s->n, 256));
            pdl(k_argmax_list, lb, 256, 0, s)(s->counters, s->worklist, s->score, s->block_best, nullptr, 0.0, 0.0, nullptr, nullptr); s->launches++;
            s->pick.block_best = s->block_best;
            s->pick.nblocks = lb;
        } else {
            // stage A: a spread sample of the most promising rows -- the maximum of the bound within each of
            // 2 x #SM strided subsets of the pool -- is scored first, one block per row.  (Taking the global
            // top rows by bound instead is worse: they cluster around the previous pick, whose neighbours have
            // just lost their gain; measured 882 vs 399 rows left for stage B at t = 3 on SYN-1M.)
            if (s->stage_a_ready_t != s->t) rc = stage_a(s, floor_score, s->ahead_cols);  // else: done during the last pass
            if (rc) return rc;
            if (s->ahead_cols) rc = launch_catchup(s, (int64_t)s->num_sms * 24, true);
            if (rc) return rc;
            rc = launch_eval(s, (int64_t)s->num_sms * 24, false);    // 3 resident blocks per SM (80 registers)
            if (rc) return rc;
            s->pick.count = s->counters;            // final argmax over the scored rows: done by k_record
            s->pick.list = s->worklist;
            s->pick.score = s->score;
        }
        CU(cudaGetLastError());
    }
    s->step_nodes[s->t] = (double)s->n_nodes;
    s->proposals = s->t + 1;
    s->last_exhaustive = exhaustive != 0;
    s->nodes_ready_t = s->stage_a_ready_t = -1;
    return make_record(s, -1, rec_out, commit_here, s->pick, pp);
}

// np.argmax over `n_records` proposals in device memory + append; with `extend` the streaming pass follows.
int commit_dev(ital_shard* s, const double* recs_dev, int n_records, int extend, bool picked = false,
               PeerWait pw = PeerWait(), int64_t rec_stride = 0) {
    if (s->t >= kMaxBatch) return fail(ITAL_EINVAL, "batches of more than %d samples are not supported", kMaxBatch);
    if (!picked) {
    s->sel_marked = true;
    pdl(k_pick_winner, 1, 256, 0, s)(recs_dev, n_records, rec_stride > 0 ? rec_stride : record_doubles(s),
                                            record_doubles(s), s->t, s->W, s->rec_in_dev,
                                            s->base_m_dev, s->base_L_dev, s->sel_dev, s->rec_hist, s->mask,
                                            s->row_offset, s->n, kSelected, pw); s->launches++;
    CU(cudaGetLastError());
    }
    if (extend && !s->lazy_rows) {      // lazy rows: the projection is extended on demand by k_catchup instead
        const int col = s->W + s->t;
        // the next step's nodes and stage-A list depend on the committed batch only, not on the pass: side stream
        const bool ahead = s->overlap && s->side && s->label_prob >= 1.0 && s->estimation == 0 && s->t + 1 <= 3 && !s->sub_mode;
        int rc = ITAL_OK;
        if (ahead) {
            CU(cudaEventRecord(s->ev_commit, s->stream));
            CU(cudaStreamWaitEvent(s->side, s->ev_commit, 0));
            cudaStream_t main_stream = s->stream;
            s->stream = s->side;
            s->t += 1;
            rc = prepare_nodes(s);
            if (rc == ITAL_OK) {
                s->nodes_ready_t = s->t;
                if (!s->last_exhaustive) {
                    rc = stage_a(s, -std::numeric_limits<double>::infinity(), true);
                    if (rc == ITAL_OK) s->stage_a_ready_t = s->t;
                }
            }
            s->t -= 1;
            s->stream = main_stream;
            if (rc) return rc;
            CU(cudaGetLastError());
            CU(cudaEventRecord(s->ev_side, s->side));
        }
        s->reserve_sm = ahead;
        rc = s->x_dtype == ITAL_F32 ? launch_extend_t<float>(s, col, 0, 0.0, 0)
                                    : launch_extend_t<double>(s, col, 0, 0.0, 0);
        s->reserve_sm = false;
        if (rc) return rc;
        if (ahead) CU(cudaStreamWaitEvent(s->stream, s->ev_side, 0));
    }
    s->t += 1;
    return ITAL_OK;
}


// ---- the fused persistent fetch kernel (ital_fused.cuh) ------------------------------------------------------------
int fused_chunk_cap(const ital_shard* s) {
    const int64_t cap = snq::capacity_for(kFusedMaxSteps - 1);
    int c = (int)((cap + s->num_sms - 1) / s->num_sms);
    return (c + 1) & ~1;
}

// Can the running fetch start with k_fetch_fused?  Users who label every sample, pruned, projections on demand, and
// the shapes the kernel's shared-memory staging covers.
bool fused_applies(const ital_shard* s, int exhaustive, bool peer) {
    if (!s->fused || !s->lazy_rows || exhaustive || s->label_prob < 1.0 || s->estimation != 0) return false;
    const int C = fused_chunk_cap(s);
    if (C > kFusedThreads) return false;
    const FusedSmem L(record_doubles(s), s->w_cap, C, s->num_sms);
    if (L.total * sizeof(double) > 200 * 1024) return false;
    if (peer && s->xg_world > kFusedThreads) return false;
    return true;
}

// Steps 0 .. steps-1 of the greedy loop in one cooperative launch.  Nothing here waits for the GPU.
int launch_fused(ital_shard* s, int steps, bool more_follow, bool peer) {
    int rc = ensure_nodes(s, snq::capacity_for(kFusedMaxSteps - 1));
    if (rc) return rc;
    FusedArgs a = {};
    a.X = s->X;
    a.n = s->n;
    a.d = (int)s->d;
    a.d_pad = (int)s->d_pad;
    a.row_offset = s->row_offset;
    a.sqn = s->sqn;
    a.m = s->m;
    a.v = s->v;
    a.U = s->U;
    a.ldu = s->ldu;
    a.mask = s->mask;
    a.mask_rw = more_follow ? s->mask : nullptr;
    a.sel_bits = kSelected;
    a.gain = s->gain;
    a.score = s->score;
    a.tags = s->tags;
    a.epoch = s->epoch;
    a.W = s->W;
    a.w_cap = s->w_cap;
    a.k = steps;
    a.var = s->var;
    a.neg2ls2 = -2.0 * (s->ls * s->ls);
    a.log1p_eps = s->log1p_eps;
    a.flag_var = 100.0 * s->noise;
    a.margin = kPruneMargin;
    const int t_saved = s->t;
    for (int t = 0; t < kFusedMaxSteps; ++t) {
        s->t = t;
        a.shift_coef[t] = step_shift_coef(s);
        a.order[t] = snq::order_for(t);
        a.node_cap[t] = snq::capacity_for(t);
    }
    s->t = t_saved;
    a.gl_x = s->gl_dev;
    a.gl_w = s->gl_dev + (snq::kMaxOrder + 1) * 64;
    a.phi = s->phi_dev;
    a.htab = s->htab_dev;
    a.R = snq::kR;
    a.w_min = snq::kWMin;
    a.q_min = snq::kQMin;
    a.chunk_cap = fused_chunk_cap(s);
    a.nodes4 = s->eta_dev;
    a.group_begin = s->group_dev;
    a.masses = s->masses_dev;
    a.hbase = s->hbase_dev;
    a.rec_hist = s->rec_hist;
    a.rec_len = record_doubles(s);
    a.rec_in = s->rec_in_dev;
    a.base_m = s->base_m_dev;
    a.base_L = s->base_L_dev;
    a.sel = s->sel_dev;
    a.stats = s->stats_dev;
    a.counters = s->counters;
    a.blk_best = s->f_best;
    a.blk_cnt = s->f_cnt;
    a.blk_mass = s->f_mass;
    a.stage_rows = s->f_stage;
    a.worklist = s->worklist;
    a.barrier = s->f_bar;
    a.bar_base = s->f_bar_count;
    a.want_scores = 0;
    a.trace = s->f_trace_on ? s->f_trace : nullptr;
    if (peer) {
        a.pp.peer_base = s->xg_peer_dev;
        a.pp.G = s->xg_world;
        a.pp.rank = s->xg_rank;
        a.pp.slot_doubles = s->xg_slot;
        a.pp.epoch = s->xg_epoch + 1;
        a.flags = reinterpret_cast<const unsigned long long*>(s->xg_local);
        a.slots = reinterpret_cast<const double*>(s->xg_local + 256);
        a.peer_error = s->xg_error_dev;
        s->xg_epoch += (unsigned long long)steps;
    }
    const int grid = s->num_sms;
    const FusedSmem L(a.rec_len, a.w_cap, a.chunk_cap, grid);
    const size_t smem = L.total * sizeof(double);
    void* params[] = {&a};
    const void* fn = s->x_dtype == ITAL_F32 ? (const void*)k_fetch_fused<float> : (const void*)k_fetch_fused<double>;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kFusedThreads), params, smem, s->stream));
    s->launches++;
    s->f_bar_count += (unsigned)grid * (unsigned)(1 + 5 * (steps - 1));
    if (more_follow) s->sel_marked = true;
    s->step_nodes[0] = 1.0;
    for (int t = 1; t < steps; ++t) s->step_nodes[t] = (double)snq::capacity_for(t);
    s->proposals = steps;
    s->fused_steps_last = steps;
    return ITAL_OK;
}

void peer_close(ital_shard* s) {
    for (int g = 0; g < (int)s->xg_peer.size(); ++g)
        if (g != s->xg_rank && s->xg_peer[g]) cudaIpcCloseMemHandle(s->xg_peer[g]);
    s->xg_peer.clear();
    s->xg_ready = false;
}

void free_all(ital_shard* s) {
    cudaSetDevice(s->device);
    if (s->side) cudaStreamDestroy(s->side);
    if (s->ev_commit) cudaEventDestroy(s->ev_commit);
    if (s->ev_side) cudaEventDestroy(s->ev_side);
    for (int k = 0; k < 16; ++k) {
        if (s->ev_win[k]) cudaEventDestroy(s->ev_win[k]);
    }
    peer_close(s);
    if (s->xg_local) cudaFree(s->xg_local);
    if (s->xg_peer_dev) cudaFree(s->xg_peer_dev);
    if (s->xg_error_dev) cudaFree(s->xg_error_dev);
    void* ptrs[] = {s->X, s->sqn, s->m, s->v, s->U, s->gain, s->score, s->mask, s->worklist, s->counters,
                    s->block_best, s->best, s->thr_dev, s->rec_dev, s->rec_in_dev, s->idx_dev, s->eta_dev,
                    s->w_dev, s->masses_dev, s->group_dev, s->orth_dev, s->eta_raw, s->w_raw, s->orth_raw, s->gl_dev, s->phi_dev, s->htab_dev, s->base_m_dev, s->base_L_dev, s->sel_dev,
                    s->hbase_dev, s->tags, s->f_best, s->f_cnt, s->f_stage, s->f_mass, s->f_bar, s->f_trace, s->rec_hist, s->mext_dev, s->g_eta, s->g_w, s->g_mass, s->g_begin, s->g_set0, s->g_lut, s->sb_eta, s->sb_w, s->sb_small, s->sb_begin, s->sc_dbl, s->sc_int, s->tn_int, s->clip_sub, s->clip_int, s->clip_dbl, s->clip_list2, s->clip_desc, s->clip_eta, s->clip_w, s->lab_x_dev, s->lab_sqn_dev, s->w_vec_dev, s->LK_dev, s->beta_dev, s->upd_idx_dev, s->upd_y_dev,
                    s->sort_keys, s->sort_rows, s->sort_hist, s->sort_out_idx, s->sort_out_val};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (s->rec_host) cudaFreeHost(s->rec_host);
    if (s->rec_in_host) cudaFreeHost(s->rec_in_host);
    if (s->sel_host) cudaFreeHost(s->sel_host);
    if (s->mext_host) cudaFreeHost(s->mext_host);
}

int reset_model(ital_shard* s) {
    s->W = 0;
    s->t = 0;
    s->fetching = false;
    s->w_valid = false;
    const int blocks = grid_for(s, s->n, 256);
    pdl(k_fill, blocks, 256, 0, s)(s->m, s->n, 0.0); s->launches++;
    pdl(k_fill, blocks, 256, 0, s)(s->v, s->n, s->var); s->launches++;
    CU(cudaGetLastError());
    // candidates are the pool rows only (queries sit behind them, retrieval_base.py:40,84)
    std::vector<uint8_t> mk((size_t)s->n, 0);
    for (int64_t i = 0; i < s->n; ++i)
        if (s->row_offset + i >= s->n_data) mk[i] = kNotCandidate;
    CU(copy_async(s, s->mask, mk.data(), (size_t)s->n, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return ITAL_OK;
}

}  // namespace

extern "C" {

const char* ital_last_error(void) { return g_err.c_str(); }
int ital_version(void) { return 100; }

int ital_create(ital_shard** out, int device, const void* X, int x_dtype, int64_t n_local, int64_t d,
                int64_t row_offset, int64_t n_data, double length_scale, double var, double noise) {
    if (!out || !X || n_local <= 0 || d <= 0 || (x_dtype != ITAL_F32 && x_dtype != ITAL_F64))
        return fail(ITAL_EINVAL, "ital_create: bad arguments");
    if (n_local >= (int64_t)1 << 31) return fail(ITAL_EINVAL, "ital_create: at most 2^31-1 rows per shard");
    if (!(length_scale > 0)) return fail(ITAL_EINVAL, "ital_create: length_scale must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ITAL_ECUDA, "ital_create: no CUDA device (%s); this path has no CPU fallback",
                    cudaGetErrorString(e));
    CU(cudaSetDevice(device));
    ital_shard* s = new ital_shard();
    s->device = device;
    s->x_dtype = x_dtype;
    s->n = n_local;
    s->d = d;
    const int64_t esize = x_dtype == ITAL_F32 ? 4 : 8;
    const int64_t row_elems = 512 / esize;                       // rows padded to a multiple of 512 bytes
    s->d_pad = (d + row_elems - 1) / row_elems * row_elems;
    s->row_offset = row_offset;
    s->n_data = n_data;
    if (const char* env = std::getenv("ITAL_B200_PDL")) s->pdl = env[0] != '0';    // (A/B comparisons)
    if (const char* env = std::getenv("ITAL_B200_OVERLAP")) s->overlap = env[0] != '0';
    if (const char* env = std::getenv("ITAL_B200_FUSED")) s->fused = env[0] != '0';
    if (const char* env = std::getenv("ITAL_B200_DEVICE_LATTICE")) s->device_lattice = env[0] != '0';
    if (const char* env = std::getenv("ITAL_B200_RESERVE")) s->reserve_sms = std::max(1, std::min(64, atoi(env)));
    s->ldu = (n_local + 31) / 32 * 32;
    s->ls = length_scale;
    s->var = var;
    s->noise = noise;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    s->num_sms = prop.multiProcessorCount;
    int rc = ITAL_OK;
    auto body = [&]() -> int {
        CU(cudaMalloc(&s->X, (size_t)s->n * s->d_pad * esize));
        if (s->d_pad == d) {
            CU(copy_sync(s, s->X, X, (size_t)s->n * d * esize, cudaMemcpyHostToDevice));
        } else {
            CU(cudaMemset(s->X, 0, (size_t)s->n * s->d_pad * esize));
            CU(cudaMemcpy2D(s->X, (size_t)s->d_pad * esize, X, (size_t)d * esize, (size_t)d * esize, (size_t)s->n,
                            cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&s->sqn, (size_t)s->n * sizeof(double)));
        CU(cudaMalloc(&s->m, (size_t)s->n * sizeof(double)));
        CU(cudaMalloc(&s->v, (size_t)s->n * sizeof(double)));
        CU(cudaMalloc(&s->gain, (size_t)s->n * sizeof(double)));
        CU(cudaMalloc(&s->score, (size_t)s->n * sizeof(double)));
        CU(cudaMemset(s->gain, 0, (size_t)s->n * sizeof(double)));      // (read by the worklist scans before the first score)
        CU(cudaMemset(s->score, 0, (size_t)s->n * sizeof(double)));
        CU(cudaMalloc(&s->mask, (size_t)s->n));
        CU(cudaMalloc(&s->tags, (size_t)s->n * sizeof(uint32_t)));
        CU(cudaMemset(s->tags, 0, (size_t)s->n * sizeof(uint32_t)));
        CU(cudaMalloc(&s->f_best, (size_t)s->num_sms * kFusedTeams * sizeof(Best)));
        CU(cudaMalloc(&s->f_cnt, (size_t)s->num_sms * 8 * sizeof(int)));
        CU(cudaMalloc(&s->f_stage, (size_t)s->num_sms * kFusedTeams * sizeof(int)));
        CU(cudaMalloc(&s->f_mass, (size_t)s->num_sms * 8 * sizeof(double)));
        CU(cudaMalloc(&s->f_bar, sizeof(unsigned)));
        CU(cudaMemset(s->f_bar, 0, sizeof(unsigned)));
        CU(cudaMalloc(&s->worklist, (size_t)s->n * sizeof(int)));
        CU(cudaMalloc(&s->counters, 8 * sizeof(int)));     // [4] ticket of k_argmax_rows
        CU(cudaMemset(s->counters, 0, 8 * sizeof(int)));
        CU(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->ev_commit, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_side, cudaEventDisableTiming));
        for (int k = 0; k < 16; ++k) {
            CU(cudaEventCreateWithFlags(&s->ev_win[k], cudaEventDisableTiming));
        }
        CU(cudaMalloc(&s->block_best, kArgmaxBlocks * sizeof(Best)));
        CU(cudaMalloc(&s->best, 2 * sizeof(Best)));
        CU(cudaMalloc(&s->thr_dev, sizeof(double)));
        CU(cudaMalloc(&s->base_m_dev, 16 * sizeof(double)));
        CU(cudaMalloc(&s->base_L_dev, 16 * 16 * sizeof(double)));
        CU(cudaMemset(s->base_m_dev, 0, 16 * sizeof(double)));             // (copied whole to the host by the paths that
        CU(cudaMemset(s->base_L_dev, 0, 16 * 16 * sizeof(double)));        //  build node sets there)
        CU(cudaMalloc(&s->sel_dev, 32 * sizeof(double) + 16 * 4 * sizeof(int)));   // selection list, then the step stats
        CU(cudaMalloc(&s->hbase_dev, 2 * sizeof(double)));      // [0] H(base), [1] total quadrature mass
        CU(cudaMemset(s->hbase_dev, 0, 2 * sizeof(double)));
        CU(cudaMalloc(&s->masses_dev, 1024 * sizeof(double)));
        CU(cudaMalloc(&s->group_dev, 1025 * sizeof(int)));
        s->stats_dev = reinterpret_cast<int*>(s->sel_dev + 32);
        CU(cudaMallocHost(&s->sel_host, 32 * sizeof(double) + 16 * 4 * sizeof(int)));
        s->stats_host = reinterpret_cast<int*>(s->sel_host + 32);
        {   // Taylor coefficients of the standard normal CDF around the grid points of phi_tab
            std::vector<double> tab = build_phi_table();
            CU(cudaMalloc(&s->phi_dev, tab.size() * sizeof(double)));
            CU(copy_sync(s, s->phi_dev, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        {   // polynomial table of the closed-form first step (h_tab)
            std::vector<double> tab = build_h_table();
            CU(cudaMalloc(&s->htab_dev, tab.size() * sizeof(double)));
            CU(copy_sync(s, s->htab_dev, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        {   // Gauss-Legendre tables for every order the panel split can ask for
            const snq::GaussLegendre& G = snq::gl();
            std::vector<double> tab((size_t)2 * (snq::kMaxOrder + 1) * 64, 0.0);
            for (int nn = 1; nn <= snq::kMaxOrder; ++nn)
                for (int i = 0; i < nn; ++i) {
                    tab[(size_t)nn * 64 + i] = G.x[nn][i];
                    tab[(size_t)(snq::kMaxOrder + 1) * 64 + (size_t)nn * 64 + i] = G.w[nn][i];
                }
            CU(cudaMalloc(&s->gl_dev, tab.size() * sizeof(double)));
            CU(copy_sync(s, s->gl_dev, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        int r = ensure_width(s, 32);
        if (r) return r;
        const int blocks = grid_for(s, s->n, 8);
        if (x_dtype == ITAL_F32)
            pdl(k_sqnorm<float>, blocks, 256, 0, s)((const float*)s->X, s->n, (int)s->d_pad, s->sqn);
        else
            pdl(k_sqnorm<double>, blocks, 256, 0, s)((const double*)s->X, s->n, (int)s->d_pad, s->sqn); s->launches++;
        CU(cudaGetLastError());
        return reset_model(s);
    };
    rc = body();
    if (rc != ITAL_OK) {
        free_all(s);
        delete s;
        return rc;
    }
    *out = s;
    return ITAL_OK;
}

int ital_destroy(ital_shard* s) {
    if (!s) return ITAL_OK;
    free_all(s);
    delete s;
    return ITAL_OK;
}

int ital_set_stream(ital_shard* s, void* cuda_stream) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    s->stream = (cudaStream_t)cuda_stream;
    return ITAL_OK;
}

int ital_reset(ital_shard* s) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    CU(cudaSetDevice(s->device));
    return reset_model(s);
}

int64_t ital_record_doubles(const ital_shard* s) { return s ? record_doubles(s) : 0; }
int64_t ital_width_cap(const ital_shard* s) { return s ? s->w_cap : 0; }
int64_t ital_width(const ital_shard* s) { return s ? s->W : 0; }

int ital_export_points(ital_shard* s, int q, const int64_t* global_idx, double* records) {
    if (!s || q < 0 || (q > 0 && (!global_idx || !records))) return fail(ITAL_EINVAL, "ital_export_points: bad arguments");
    CU(cudaSetDevice(s->device));
    int rc = ensure_record_buffers(s, std::max(q, 1));
    if (rc) return rc;
    const int64_t rl = record_doubles(s);
    bool any = false;
    for (int a = 0; a < q; ++a) {
        const int64_t loc = global_idx[a] - s->row_offset;
        if (loc >= 0 && loc < s->n) {
            rc = make_record(s, loc, s->rec_dev + a * rl);
            if (rc) return rc;
            any = true;
        }
    }
    if (any) {
        CU(copy_async(s, s->rec_host, s->rec_dev, (size_t)q * rl * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    for (int a = 0; a < q; ++a) {
        const int64_t loc = global_idx[a] - s->row_offset;
        if (loc >= 0 && loc < s->n) memcpy(records + a * rl, s->rec_host + a * rl, (size_t)rl * sizeof(double));
        else memset(records + a * rl, 0, (size_t)rl * sizeof(double));
    }
    return ITAL_OK;
}

int ital_add_labelled(ital_shard* s, const double* record, double y) {
    if (!s || !record) return fail(ITAL_EINVAL, "ital_add_labelled: bad arguments");
    return ital_add_labelled_many(s, 1, record, &y);
}

// GaussianProcess.update with several samples (ital/gp.py:164-200): q <= 4 labelled points, whose records were all
// exported in the CURRENT state of the model, enter with ONE pass over the pool.  The q x q triangle of the block
// Cholesky extension among the new points is computed here on the host from the records (the multi-GPU path: the
// records have been summed over the shards on the host anyway; a single shard uses ital_update_labelled).
int ital_add_labelled_many(ital_shard* s, int q, const double* records, const double* y) {
    if (!s || q < 1 || !records || !y) return fail(ITAL_EINVAL, "ital_add_labelled_many: bad arguments");
    if (s->fetching) return fail(ITAL_ESTATE, "ital_add_labelled_many: a fetch is in progress");
    if (q > 4) return fail(ITAL_EINVAL, "ital_add_labelled_many: at most 4 points per pass");
    CU(cudaSetDevice(s->device));
    const int W = s->W;
    const int64_t old_cap = s->w_cap, old_rl = record_doubles(s);
    const int64_t d = s->d;
    std::vector<std::vector<double>> u(q), x(q);
    std::vector<double> hm(q), hv(q), hsq(q), hidx(q);
    for (int a = 0; a < q; ++a) {
        const double* r = records + (int64_t)a * old_rl;
        hidx[a] = r[0]; hm[a] = r[2]; hsq[a] = r[4]; hv[a] = r[5];
        u[a].assign(r + ITAL_RECORD_HEADER, r + ITAL_RECORD_HEADER + W);
        x[a].assign(r + ITAL_RECORD_HEADER + old_cap, r + ITAL_RECORD_HEADER + old_cap + d);
    }
    int rc = ensure_width(s, W + q);
    if (rc) return rc;
    if ((rc = ensure_model(s, W + q))) return rc;
    // block Cholesky extension among the new points
    const double neg2ls2 = -2.0 * (s->ls * s->ls);
    double tri[16] = {0}, piv[4] = {0}, beta[4] = {0};
    for (int a = 0; a < q; ++a) {
        double cv = hv[a], ma = hm[a];
        for (int b = 0; b < a; ++b) {
            double dot = 0.0;
            for (int64_t j = 0; j < d; ++j) dot += x[a][j] * x[b][j];
            double num = s->var * std::exp((hsq[a] + hsq[b] - 2.0 * dot) / neg2ls2);
            for (int j = 0; j < W; ++j) num -= u[a][j] * u[b][j];
            for (int c = 0; c < b; ++c) num -= tri[a * 4 + c] * tri[b * 4 + c];
            const double e = num / piv[b];
            tri[a * 4 + b] = e;
            cv -= e * e;
            ma += e * beta[b];
        }
        piv[a] = std::sqrt(std::max(cv + s->noise, 2.3e-308));
        beta[a] = (y[a] - ma) / piv[a];
    }
    // device block: MultiExt header, z[q][d_pad] (zero padded), ur[q][W]
    const size_t hdr_d = sizeof(MultiExt) / sizeof(double);
    const size_t need = hdr_d + (size_t)q * s->d_pad + (size_t)q * W;
    if ((rc = ensure_mext(s, q, W))) return rc;
    CU(cudaStreamSynchronize(s->stream));               // the pinned block may still feed the previous pass
    MultiExt* h = reinterpret_cast<MultiExt*>(s->mext_host);
    memset(s->mext_host, 0, need * sizeof(double));
    for (int a = 0; a < q; ++a) {
        h->zn[a] = hsq[a];
        h->piv[a] = piv[a];
        h->beta[a] = beta[a];
        memcpy(s->mext_host + hdr_d + (size_t)a * s->d_pad, x[a].data(), (size_t)d * sizeof(double));
        memcpy(s->mext_host + hdr_d + (size_t)q * s->d_pad + (size_t)a * W, u[a].data(), (size_t)W * sizeof(double));
    }
    memcpy(h->tri, tri, sizeof tri);
    CU(copy_async(s, s->mext_dev, s->mext_host, need * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    pdl(k_append_model, 1, 256, 0, s)(model_refs(s), q, W, (int)s->d, (int)s->d_pad, s->mext_dev); s->launches++;
    CU(cudaGetLastError());
    if (q == 1) {
        // one point: the single-column pass, fed by the point record in the current layout
        std::vector<double> rec((size_t)record_doubles(s), 0.0);
        memcpy(rec.data(), records, ITAL_RECORD_HEADER * sizeof(double));
        std::copy(u[0].begin(), u[0].end(), rec.begin() + ITAL_RECORD_HEADER);
        std::copy(x[0].begin(), x[0].end(), rec.begin() + ITAL_RECORD_HEADER + s->w_cap);
        rc = extend_with_record(s, rec.data(), W, 1, y[0], false);
    } else {
        rc = s->x_dtype == ITAL_F32 ? launch_extend_multi<float>(s, q, W) : launch_extend_multi<double>(s, q, W);
    }
    if (rc) return rc;
    std::vector<int64_t> gidx(q);
    for (int a = 0; a < q; ++a) gidx[a] = (int64_t)hidx[a];
    s->w_valid = false;
    s->W += q;
    return ital_mark_seen(s, q, gidx.data());
}

// ActiveRetrievalBase.update -> GaussianProcess.update (ital/retrieval_base.py:105-126, ital/gp.py:164-200) for a
// learner whose rows all live on this shard: q <= 4 labelled pool rows enter the model with one small kernel (records,
// triangle among the new points, model rows -- k_prepare_labelled) and ONE pass over the pool.  Nothing here waits
// for the GPU and nothing comes back to the host.
int ital_update_labelled(ital_shard* s, int q, const int64_t* global_idx, const double* y) {
    if (!s || q < 1 || q > 4 || !global_idx || !y) return fail(ITAL_EINVAL, "ital_update_labelled: bad arguments (1 <= q <= 4)");
    if (s->fetching) return fail(ITAL_ESTATE, "ital_update_labelled: a fetch is in progress");
    for (int a = 0; a < q; ++a) {
        const int64_t loc = global_idx[a] - s->row_offset;
        if (loc < 0 || loc >= s->n) return fail(ITAL_EINVAL, "ital_update_labelled: row %lld is not on this shard", (long long)global_idx[a]);
        for (int b = 0; b < a; ++b)
            if (global_idx[a] == global_idx[b]) return fail(ITAL_EINVAL, "ital_update_labelled: row %lld given twice", (long long)global_idx[a]);
    }
    CU(cudaSetDevice(s->device));
    const int W = s->W;
    int rc = ensure_width(s, W + q);
    if (rc) return rc;
    if ((rc = ensure_model(s, W + q))) return rc;
    if ((rc = ensure_mext(s, q, W))) return rc;
    if ((rc = ensure_record_buffers(s, 1))) return rc;
    if (!s->upd_idx_dev) {
        CU(cudaMalloc(&s->upd_idx_dev, 8 * sizeof(int64_t)));
        CU(cudaMalloc(&s->upd_y_dev, 8 * sizeof(double)));
    }
    // (pageable sources: the runtime stages these few bytes before returning, no wait for the GPU)
    CU(copy_async(s, s->upd_idx_dev, global_idx, (size_t)q * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream));
    CU(copy_async(s, s->upd_y_dev, y, (size_t)q * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    const double neg2ls2 = -2.0 * (s->ls * s->ls);
    if (s->x_dtype == ITAL_F32)
        pdl(k_prepare_labelled<float>, 1, 256, 0, s)(q, s->upd_idx_dev, s->upd_y_dev, s->row_offset, s->n, (const float*)s->X,
                                                     (int)s->d, (int)s->d_pad, s->sqn, s->m, s->v, s->U, s->ldu, W, s->w_cap,
                                                     s->var, neg2ls2, s->noise, s->mext_dev, s->rec_in_dev, model_refs(s),
                                                     s->mask, kSeen);
    else
        pdl(k_prepare_labelled<double>, 1, 256, 0, s)(q, s->upd_idx_dev, s->upd_y_dev, s->row_offset, s->n, (const double*)s->X,
                                                      (int)s->d, (int)s->d_pad, s->sqn, s->m, s->v, s->U, s->ldu, W, s->w_cap,
                                                      s->var, neg2ls2, s->noise, s->mext_dev, s->rec_in_dev, model_refs(s),
                                                      s->mask, kSeen);
    s->launches++;
    CU(cudaGetLastError());
    if (q == 1) {
        s->ext_rec = nullptr;       // k_extend reads rec_in_dev, written by the kernel above
        rc = s->x_dtype == ITAL_F32 ? launch_extend_t<float>(s, W, 1, y[0], 0) : launch_extend_t<double>(s, W, 1, y[0], 0);
    } else {
        rc = s->x_dtype == ITAL_F32 ? launch_extend_multi<float>(s, q, W) : launch_extend_multi<double>(s, q, W);
    }
    if (rc) return rc;
    s->w_valid = false;
    s->W += q;
    return ITAL_OK;
}

int ital_mark_seen(ital_shard* s, int64_t m, const int64_t* global_idx) {
    if (!s || m < 0 || (m > 0 && !global_idx)) return fail(ITAL_EINVAL, "ital_mark_seen: bad arguments");
    CU(cudaSetDevice(s->device));
    std::vector<int64_t> loc;
    for (int64_t k = 0; k < m; ++k) {
        const int64_t l = global_idx[k] - s->row_offset;
        if (l >= 0 && l < s->n) loc.push_back(l);
    }
    if (loc.empty()) return ITAL_OK;
    int rc = ensure_idx(s, (int64_t)loc.size());
    if (rc) return rc;
    CU(copy_async(s, s->idx_dev, loc.data(), loc.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream));
    pdl(k_mask_rows, grid_for(s, (int64_t)loc.size(), 256), 256, 0, s)(s->mask, s->idx_dev, (int64_t)loc.size(), (uint8_t)(kSeen | kUnnameable)); s->launches++;
    CU(cudaGetLastError());
    if (loc.size() * sizeof(int64_t) > 32 * 1024) CU(cudaStreamSynchronize(s->stream));   // (small pageable copies are staged)
    return ITAL_OK;
}

int ital_restrict_candidates(ital_shard* s, int64_t m, const int64_t* global_idx) {
    if (!s || (m > 0 && !global_idx)) return fail(ITAL_EINVAL, "ital_restrict_candidates: bad arguments");
    CU(cudaSetDevice(s->device));
    const int blocks = grid_for(s, s->n, 256);
    if (m < 0) {
        pdl(k_mask_all, blocks, 256, 0, s)(s->mask, s->n, (uint8_t)~kRestricted, 0); s->launches++;
        CU(cudaGetLastError());
        return ITAL_OK;
    }
    // everything restricted, then the listed rows released
    pdl(k_mask_all, blocks, 256, 0, s)(s->mask, s->n, 0xff, kRestricted); s->launches++;
    CU(cudaGetLastError());
    std::vector<int64_t> loc;
    for (int64_t k = 0; k < m; ++k) {
        const int64_t l = global_idx[k] - s->row_offset;
        if (l >= 0 && l < s->n) loc.push_back(l);
    }
    if (!loc.empty()) {
        int rc = ensure_idx(s, (int64_t)loc.size());
        if (rc) return rc;
        CU(copy_async(s, s->idx_dev, loc.data(), loc.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream));
        pdl(k_mask_clear_rows, grid_for(s, (int64_t)loc.size(), 256), 256, 0, s)(s->mask, s->idx_dev,
                                                                                       (int64_t)loc.size(), kRestricted); s->launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s->stream));
    }
    return ITAL_OK;
}

int ital_fetch_begin(ital_shard* s, double label_prob, double mistake_prob) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    if (s->W == 0) return fail(ITAL_ESTATE, "fetch before any labelled point or query (the reference fails here too: gp.K_inv is None)");
    if (!(label_prob > 0.0)) return fail(ITAL_EINVAL, "label_prob must be positive");
    if (!(mistake_prob >= 0.0 && mistake_prob <= 1.0)) return fail(ITAL_EINVAL, "mistake_prob must be in [0, 1]");
    CU(cudaSetDevice(s->device));
    if (s->fetching) {
        int rc = ital_fetch_end(s);
        if (rc) return rc;
    }
    // room for the batch's projection columns up front: the record layout stays fixed during the fetch
    int rc = ensure_width(s, s->W + kMaxBatch + 1);
    if (rc) return rc;
    rc = ensure_record_buffers(s, 1);
    if (rc) return rc;
    const int64_t hist_need = record_doubles(s) * (kMaxBatch + 1);
    if (hist_need > s->rec_hist_cap) {
        CU(cudaStreamSynchronize(s->stream));
        if (s->rec_hist) CU(cudaFree(s->rec_hist));
        s->rec_hist = nullptr;
        CU(cudaMalloc(&s->rec_hist, (size_t)hist_need * sizeof(double)));
        s->rec_hist_cap = hist_need;
    }
    // a new epoch empties every row tag (no n-sized memset per fetch); the 24-bit epoch wraps after 16M fetches
    if (++s->epoch >= (1u << 24)) {
        CU(cudaMemsetAsync(s->tags, 0, (size_t)s->n * sizeof(uint32_t), s->stream));
        s->epoch = 1;
    }
    s->sel_marked = false;
    s->hbase_seeded = false;                            // (uploaded by the first multi-kernel step; the fused kernel seeds its own)
    s->fetching = true;
    s->t = 0;
    s->nodes_ready_t = s->stage_a_ready_t = -1;
    s->proposals = 0;
    s->label_prob = label_prob;
    s->mistake_prob = mistake_prob;
    return ITAL_OK;
}

int ital_fetch_propose_dev(ital_shard* s, double floor_score, int exhaustive, double* record_dev) {
    if (!s || !record_dev) return fail(ITAL_EINVAL, "ital_fetch_propose_dev: bad arguments");
    if (!s->fetching) return fail(ITAL_ESTATE, "ital_fetch_propose_dev outside a fetch");
    CU(cudaSetDevice(s->device));
    return propose_dev(s, floor_score, exhaustive, record_dev);
}

int ital_fetch_commit_dev(ital_shard* s, const double* records_dev, int n_records, int extend) {
    if (!s || !records_dev || n_records < 1) return fail(ITAL_EINVAL, "ital_fetch_commit_dev: bad arguments");
    if (!s->fetching) return fail(ITAL_ESTATE, "ital_fetch_commit_dev outside a fetch");
    CU(cudaSetDevice(s->device));
    return commit_dev(s, records_dev, n_records, extend);
}

static int read_step_stats(ital_shard* s, int step) {
    // (after a synchronisation) diagnostics of propose number `step`
    const int* c = s->stats_host + 4 * step;
    for (double& x : s->stats) x = 0.0;
    s->stats[0] = (double)c[0];                         // rows in the final worklist
    s->stats[1] = step == 0 ? -1.0 : (double)c[2];      // rows scored by quadrature (-1: closed form for all)
    s->stats[2] = s->step_nodes[step];
    s->stats[4] = (double)c[1];
    s->stats[5] = (double)s->fused_steps_last;
    s->stats[6] = step == 0 ? 1.0 : (double)c[3];      // nodes kept (t <= 3: after the light ones are dropped)
    return ITAL_OK;
}

int ital_fetch_propose(ital_shard* s, double floor_score, int exhaustive, double* record) {
    if (!s || !record) return fail(ITAL_EINVAL, "ital_fetch_propose: bad arguments");
    if (!s->fetching) return fail(ITAL_ESTATE, "ital_fetch_propose outside a fetch");
    CU(cudaSetDevice(s->device));
    const int step = s->t;
    int rc = propose_dev(s, floor_score, exhaustive, s->rec_dev);
    if (rc) return rc;
    const int64_t rl = record_doubles(s);
    double hb = 0.0;
    CU(copy_async(s, s->rec_host, s->rec_dev, (size_t)rl * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(copy_async(s, s->stats_host, s->stats_dev, 16 * 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(copy_async(s, &hb, s->hbase_dev, sizeof(double), cudaMemcpyDeviceToHost, s->stream));   // pageable: synchronous
    CU(cudaStreamSynchronize(s->stream));
    memcpy(record, s->rec_host, (size_t)rl * sizeof(double));
    read_step_stats(s, step);
    s->stats[3] = step == 0 ? 0.0 : hb;
    return ITAL_OK;
}

int ital_variance_propose(ital_shard* s, int use_correlations, int first_pick_takes_unnameable, double* record) {
    if (!s || !record) return fail(ITAL_EINVAL, "ital_variance_propose: bad arguments");
    if (!s->fetching) return fail(ITAL_ESTATE, "ital_variance_propose outside a fetch");
    if (s->t >= kMaxBatch) return fail(ITAL_EINVAL, "batches of more than %d samples are not supported", kMaxBatch);
    CU(cudaSetDevice(s->device));
    const int blocks = std::min(kArgmaxBlocks, grid_for(s, s->n, 256));
    const int t = use_correlations ? s->t : 0;
    const uint8_t allow = (first_pick_takes_unnameable && s->t == 0) ? (uint8_t)(kSeen | kUnnameable) : (uint8_t)0;
    pdl(k_var_score, blocks, 256, 0, s)(s->n, s->v, s->U, s->ldu, s->W, t, s->base_L_dev, s->mask, allow, s->score,
                                        s->block_best); s->launches++;
    CU(cudaGetLastError());
    PickSrc pick;
    pick.block_best = s->block_best;
    pick.nblocks = blocks;
    s->proposals = s->t + 1;
    int rc = make_record(s, -1, s->rec_dev, false, pick);
    if (rc) return rc;
    const int64_t rl = record_doubles(s);
    CU(copy_async(s, s->rec_host, s->rec_dev, (size_t)rl * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(record, s->rec_host, (size_t)rl * sizeof(double));
    return ITAL_OK;
}

int ital_set_clip_cov(ital_shard* s, double clip_cov) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    if (s->fetching) return fail(ITAL_ESTATE, "ital_set_clip_cov during a fetch");
    s->clip_cov = clip_cov;
    return ITAL_OK;
}

int ital_set_sub_mode(ital_shard* s, int on) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    if (s->fetching) return fail(ITAL_ESTATE, "ital_set_sub_mode during a fetch");
    s->sub_mode = on != 0;
    return ITAL_OK;
}

int ital_fetch_propose_sub(ital_shard* s, int n_batch, int64_t only_row, double* record) {
    if (!s || !record) return fail(ITAL_EINVAL, "ital_fetch_propose_sub: bad arguments");
    if (!s->fetching || !s->sub_mode) return fail(ITAL_ESTATE, "ital_fetch_propose_sub outside a fetch in subset mode");
    if (s->lazy_rows) return fail(ITAL_ESTATE, "ital_fetch_propose_sub needs the streaming pass (lazy rows off)");
    if (!(s->label_prob >= 1.0) || s->estimation != 0)
        return fail(ITAL_EINVAL, "change_estimation_subset is built for users who label every sample (label_prob = 1) and label_estimation 'mean'");
    if (n_batch < 0 || n_batch > s->t || n_batch > 7 || s->t > kSubMaxCols)
        return fail(ITAL_EINVAL, "ital_fetch_propose_sub: %d batch samples of %d columns (at most 7 of %d)", n_batch, s->t, kSubMaxCols);
    CU(cudaSetDevice(s->device));
    int64_t only_local = -1;
    bool any_local = true;
    if (only_row >= 0) {
        only_local = only_row - s->row_offset;
        if (only_local < 0 || only_local >= s->n) { only_local = -1; any_local = false; }
    }
    const int step = std::min(s->t, 15);
    int rc = propose_sub(s, n_batch, only_local, any_local);
    if (rc) return rc;
    const int64_t rl = record_doubles(s);
    CU(copy_async(s, s->rec_host, s->rec_dev, (size_t)rl * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(copy_async(s, s->stats_host, s->stats_dev, 16 * 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(record, s->rec_host, (size_t)rl * sizeof(double));
    s->proposals = step + 1;
    read_step_stats(s, step);
    return ITAL_OK;
}

int ital_fetch_commit(ital_shard* s, const double* record) {
    if (!s || !record) return fail(ITAL_EINVAL, "ital_fetch_commit: bad arguments");
    if (!s->fetching) return fail(ITAL_ESTATE, "ital_fetch_commit outside a fetch");
    if (record[0] < 0) return fail(ITAL_EINVAL, "ital_fetch_commit: empty record");
    CU(cudaSetDevice(s->device));
    // through pinned staging into the device buffer the winner is picked from (a list of one)
    const int64_t rl = record_doubles(s);
    memcpy(s->rec_in_host, record, (size_t)rl * sizeof(double));
    CU(copy_async(s, s->rec_dev, s->rec_in_host, (size_t)rl * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    return commit_dev(s, s->rec_dev, 1, 1);
}

int ital_fetch_result(ital_shard* s, int max_out, int64_t* out_idx, double* out_scores) {
    if (!s || max_out < 0 || (max_out > 0 && !out_idx)) return fail(ITAL_EINVAL, "ital_fetch_result: bad arguments");
    CU(cudaSetDevice(s->device));
    CU(copy_async(s, s->sel_host, s->sel_dev, 32 * sizeof(double) + 16 * 4 * sizeof(int), cudaMemcpyDeviceToHost,
                  s->stream));                          // selection list and step stats in one copy
    CU(cudaStreamSynchronize(s->stream));
    int got = 0;
    for (int k = 0; k < s->t && k < max_out; ++k) {
        if (s->sel_host[2 * k] < 0) break;      // no candidate was left at that step
        out_idx[got] = (int64_t)s->sel_host[2 * k];
        if (out_scores) out_scores[got] = s->sel_host[2 * k + 1];
        ++got;
    }
    if (s->proposals > 0) read_step_stats(s, s->proposals - 1);
    return got;
}

int ital_fetch_end(ital_shard* s) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    CU(cudaSetDevice(s->device));
    if (s->fetching && s->sel_marked) {
        pdl(k_mask_all, grid_for(s, s->n, 256), 256, 0, s)(s->mask, s->n, (uint8_t)~kSelected, 0); s->launches++;
        CU(cudaGetLastError());
    }
    s->sel_marked = false;
    s->fetching = false;
    s->t = 0;
    return ITAL_OK;
}

// The greedy loop of ital_fetch / ital_fetch_peer.  Streaming mode, perfect user, pruned: PIPELINED -- the passes
// E_0, E_1, ... run back to back on the main stream, each on all SMs but `reserve_sms`; the scoring chain
// S_1, S_2, S_3 (nodes, stage A, stage B, record / winner of each step) runs on the side stream beside them.
// S_{t+1} needs the batch committed by S_t and the batch columns of the few hundred rows it scores, which k_catchup
// computes on demand from the winner records with the same bits as the passes; E_t needs only the record committed
// by S_t (read from the per-step record history, which is never overwritten within a fetch).  The critical path of
// a fetch is then S_0 + E_0 + ... + E_{k-2}: the HBM passes themselves.
int greedy_loop(ital_shard* s, int k, int exhaustive, bool peer) {
    const double ninf = -std::numeric_limits<double>::infinity();
    const bool pipelined = s->overlap && s->side && !s->lazy_rows && s->label_prob >= 1.0 && !exhaustive && s->estimation == 0;
    const int64_t rl = record_doubles(s);
    const int G = s->xg_world;
    cudaStream_t main_stream = s->stream;
    int rc = ITAL_OK;
    bool last_on_side = false;
    int it0 = 0;
    s->fused_steps_last = 0;
    if (k > 0 && fused_applies(s, exhaustive, peer)) {
        // the first (up to four) greedy steps in one persistent kernel; later steps continue below
        it0 = std::min(k, kFusedMaxSteps);
        rc = launch_fused(s, it0, k > it0, peer);
        if (rc) return rc;
        s->t = it0;
    }
    for (int it = it0; it < k && rc == ITAL_OK; ++it) {
        const bool on_side = pipelined && it >= 1 && it <= 3;
        if (on_side) {
            if (it == 1) CU(cudaStreamWaitEvent(s->side, s->ev_win[0], 0));     // (later steps follow on the side stream)
            s->stream = s->side;
            s->ahead_cols = true;
        }
        if (peer) {
            const unsigned long long epoch = ++s->xg_epoch;
            PeerPut pp;
            pp.peer_base = s->xg_peer_dev;
            pp.G = G;
            pp.rank = s->xg_rank;
            pp.slot_doubles = s->xg_slot;
            pp.epoch = epoch;
            rc = propose_dev(s, ninf, exhaustive, s->rec_dev, false, pp);
            if (rc == ITAL_OK) {
                PeerWait pw;
                pw.flags = reinterpret_cast<const unsigned long long*>(s->xg_local);
                pw.epoch = epoch;
                pw.error = s->xg_error_dev;
                const double* slots = reinterpret_cast<const double*>(s->xg_local + 256) +
                                      (int64_t)(epoch & 1) * G * s->xg_slot;     // slots xg_slot doubles apart
                s->sel_marked = true;
                pdl(k_pick_winner, 1, 256, 0, s)(slots, G, s->xg_slot, rl, s->t, s->W, s->rec_in_dev, s->base_m_dev,
                                                 s->base_L_dev, s->sel_dev, s->rec_hist, s->mask, s->row_offset, s->n,
                                                 kSelected, pw); s->launches++;
            }
        } else {
            // single shard: the record kernel commits the winner itself (no separate pick)
            rc = propose_dev(s, ninf, exhaustive, s->rec_dev, true);
        }
        cudaError_t e = cudaGetLastError();
        if (rc == ITAL_OK && e == cudaSuccess && pipelined) e = cudaEventRecord(s->ev_win[it], s->stream);
        s->stream = main_stream;
        s->ahead_cols = false;
        if (rc) break;
        CU(e);
        last_on_side = on_side;
        if (it + 1 < k && !s->lazy_rows) {      // lazy rows: the projection is extended on demand by k_catchup instead
            if (on_side) CU(cudaStreamWaitEvent(main_stream, s->ev_win[it], 0));
            s->reserve_sm = pipelined && it + 1 <= 3;           // the next step is scored beside this pass
            s->ext_rec = pipelined ? s->rec_hist + (int64_t)s->t * rl : nullptr;
            rc = s->x_dtype == ITAL_F32 ? launch_extend_t<float>(s, s->W + s->t, 0, 0.0, 0)
                                        : launch_extend_t<double>(s, s->W + s->t, 0, 0.0, 0);
            s->reserve_sm = false;
            s->ext_rec = nullptr;
            if (rc) break;
        }
        s->t += 1;
    }
    if (rc == ITAL_OK && last_on_side) CU(cudaStreamWaitEvent(main_stream, s->ev_win[k - 1], 0));
    if (rc != ITAL_OK && pipelined) cudaStreamSynchronize(s->side);
    return rc;
}

int ital_peer_export(ital_shard* s, int world, int rank, void* handle_out, int64_t handle_bytes) {
    if (!s || world < 2 || world > 32 || rank < 0 || rank >= world || !handle_out ||
        handle_bytes < (int64_t)sizeof(cudaIpcMemHandle_t))
        return fail(ITAL_EINVAL, "ital_peer_export: bad arguments");
    CU(cudaSetDevice(s->device));
    peer_close(s);
    if (s->xg_local) { CU(cudaFree(s->xg_local)); s->xg_local = nullptr; }
    s->xg_world = world;
    s->xg_rank = rank;
    s->xg_slot = ITAL_RECORD_HEADER + 2048 + s->d;          // records with up to 2048 projection entries
    const size_t bytes = 256 + (size_t)2 * world * s->xg_slot * sizeof(double);
    CU(cudaMalloc(&s->xg_local, bytes));
    CU(cudaMemset(s->xg_local, 0, bytes));
    if (!s->xg_error_dev) CU(cudaMalloc(&s->xg_error_dev, sizeof(int)));
    CU(cudaMemset(s->xg_error_dev, 0, sizeof(int)));
    CU(cudaDeviceSynchronize());
    s->xg_epoch = 0;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->xg_local));
    memset(handle_out, 0, (size_t)handle_bytes);
    memcpy(handle_out, &h, sizeof h);
    return ITAL_OK;
}

int ital_peer_connect(ital_shard* s, const void* handles, int64_t handle_bytes) {
    if (!s || !handles || !s->xg_local || handle_bytes < (int64_t)sizeof(cudaIpcMemHandle_t))
        return fail(ITAL_EINVAL, "ital_peer_connect: bad arguments (ital_peer_export first)");
    CU(cudaSetDevice(s->device));
    peer_close(s);
    s->xg_peer.assign((size_t)s->xg_world, nullptr);
    for (int g = 0; g < s->xg_world; ++g) {
        if (g == s->xg_rank) { s->xg_peer[g] = s->xg_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)handles + (size_t)g * handle_bytes, sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            peer_close(s);
            return fail(ITAL_ECUDA, "ital_peer_connect: cannot map the exchange buffer of shard %d (%s)", g,
                        cudaGetErrorString(e));
        }
        s->xg_peer[g] = (unsigned char*)p;
    }
    if (!s->xg_peer_dev) CU(cudaMalloc(&s->xg_peer_dev, 32 * sizeof(unsigned char*)));
    CU(copy_sync(s, s->xg_peer_dev, s->xg_peer.data(), (size_t)s->xg_world * sizeof(unsigned char*), cudaMemcpyHostToDevice));
    s->xg_ready = true;
    return ITAL_OK;
}

int ital_peer_disconnect(ital_shard* s) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    peer_close(s);
    return ITAL_OK;
}

int64_t ital_peer_slot_doubles(const ital_shard* s) { return s && s->xg_ready ? s->xg_slot : 0; }

int ital_fetch_peer(ital_shard* s, int k, double label_prob, double mistake_prob, int exhaustive, int64_t* out_idx,
                    double* out_scores) {
    if (!s || k < 0 || (k > 0 && !out_idx)) return fail(ITAL_EINVAL, "ital_fetch_peer: bad arguments");
    if (!s->xg_ready) return fail(ITAL_ESTATE, "ital_fetch_peer: no peer exchange (ital_peer_export / ital_peer_connect)");
    if (k > kMaxBatch) return fail(ITAL_EINVAL, "batches of more than %d samples are not supported", kMaxBatch);
    if ((label_prob < 1.0 || s->estimation != 0) && k > kMaxBatchGeneral)
        return fail(ITAL_EINVAL, "label_prob < 1 (and label_estimation other than 'mean') supports batches of at most %d samples", kMaxBatchGeneral);
    int rc = ital_fetch_begin(s, label_prob, mistake_prob);
    if (rc) return rc;
    const int64_t rl = record_doubles(s);
    if (rl > s->xg_slot) {
        ital_fetch_end(s);
        return fail(ITAL_ESTATE, "ital_fetch_peer: records of %lld doubles exceed the exchange slots", (long long)rl);
    }
    // every shard enqueues the same k steps; nothing waits for the host, the shards meet in k_pick_winner
    rc = greedy_loop(s, k, exhaustive, true);
    int got = 0;
    if (rc == ITAL_OK) {
        got = ital_fetch_result(s, k, out_idx, out_scores);
        if (got >= 0) {
            int err = 0;
            CU(copy_sync(s, &err, s->xg_error_dev, sizeof err, cudaMemcpyDeviceToHost));
            if (err) {
                CU(cudaMemset(s->xg_error_dev, 0, sizeof(int)));
                rc = fail(ITAL_ECUDA, "ital_fetch_peer: a shard did not deliver its proposal within 5 s");
            }
        } else {
            rc = got;
        }
    }
    int rc2 = ital_fetch_end(s);
    if (rc) return rc;
    if (rc2) return rc2;
    return got;
}

int ital_fetch(ital_shard* s, int k, double label_prob, double mistake_prob, int exhaustive, int64_t* out_idx,
               double* out_scores) {
    if (!s || k < 0 || (k > 0 && !out_idx)) return fail(ITAL_EINVAL, "ital_fetch: bad arguments");
    if (k > kMaxBatch) return fail(ITAL_EINVAL, "batches of more than %d samples are not supported", kMaxBatch);
    if ((label_prob < 1.0 || s->estimation != 0) && k > kMaxBatchGeneral)
        return fail(ITAL_EINVAL, "label_prob < 1 (and label_estimation other than 'mean') supports batches of at most %d samples", kMaxBatchGeneral);
    int rc = ital_fetch_begin(s, label_prob, mistake_prob);
    if (rc) return rc;
    // the whole greedy loop is enqueued without waiting for the GPU; one read-back at the end
    rc = greedy_loop(s, k, exhaustive, false);
    int got = 0;
    if (rc == ITAL_OK) {
        got = ital_fetch_result(s, k, out_idx, out_scores);
        if (got < 0) rc = got;
    }
    int rc2 = ital_fetch_end(s);
    if (rc) return rc;
    if (rc2) return rc2;
    return got;
}

int ital_fetch_stats(const ital_shard* s, double* out8) {
    if (!s || !out8) return fail(ITAL_EINVAL, "ital_fetch_stats: bad arguments");
    memcpy(out8, s->stats, sizeof s->stats);
    return ITAL_OK;
}

static int copy_vec(ital_shard* s, const double* dev, double* out) {
    CU(cudaSetDevice(s->device));
    CU(copy_async(s, out, dev, (size_t)s->n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return ITAL_OK;
}

int ital_last_scores(ital_shard* s, double* out) {
    if (!s || !out) return fail(ITAL_EINVAL, "bad arguments");
    int rc = copy_vec(s, s->score, out);
    if (rc) return rc;
    if (s->proposals > 1) {      // steps after the first: only rows stamped with the step were scored in it
        std::vector<uint32_t> st((size_t)s->n);
        CU(copy_sync(s, st.data(), s->tags, (size_t)s->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        const uint32_t want = (s->epoch << 8) | (uint32_t)(s->proposals - 1);
        for (int64_t i = 0; i < s->n; ++i)
            if ((st[i] & 0xffffff0fu) != want) out[i] = std::numeric_limits<double>::quiet_NaN();
    }
    if (s->mistake_prob > 0.0 && s->proposals > 0) {    // same additive constant as the records carry
        double hb[2];
        CU(copy_sync(s, hb, s->hbase_dev, sizeof hb, cudaMemcpyDeviceToHost));
        const int t_saved = s->t;
        s->t = s->proposals - 1;
        const double shift = step_shift_coef(s) * hb[1];
        s->t = t_saved;
        for (int64_t i = 0; i < s->n; ++i)
            if (out[i] == out[i]) out[i] += shift;
    }
    return ITAL_OK;
}
int ital_rel_mean(ital_shard* s, double* out) {
    if (!s || !out) return fail(ITAL_EINVAL, "bad arguments");
    return copy_vec(s, s->m, out);
}
int ital_rel_var(ital_shard* s, double* out) {
    if (!s || !out) return fail(ITAL_EINVAL, "bad arguments");
    return copy_vec(s, s->v, out);
}

// stable LSD radix sort of the local pool rows by descending mean (masked rows last if `masked`); the sorted rows end
// up in sort_rows[0 .. np)
static int sort_rows_by_mean(ital_shard* s, int64_t np, bool masked) {
    if (np >= ((int64_t)1 << 32)) return fail(ITAL_EINVAL, "more than 2^32 rows in one shard");
    const int tiles = (int)((np + kSortTile - 1) / kSortTile);
    if (np > s->sort_cap) {
        for (void** p : {(void**)&s->sort_keys, (void**)&s->sort_rows, (void**)&s->sort_hist, (void**)&s->sort_out_idx,
                         (void**)&s->sort_out_val})
            if (*p) { CU(cudaFree(*p)); *p = nullptr; }
        s->sort_cap = 0;
        CU(cudaMalloc(&s->sort_keys, 2 * (size_t)np * sizeof(uint64_t)));
        CU(cudaMalloc(&s->sort_rows, 2 * (size_t)np * sizeof(uint32_t)));
        CU(cudaMalloc(&s->sort_hist, ((size_t)256 * tiles + 256) * sizeof(uint32_t)));
        CU(cudaMalloc(&s->sort_out_idx, (size_t)np * sizeof(int64_t)));
        CU(cudaMalloc(&s->sort_out_val, (size_t)np * sizeof(double)));
        s->sort_cap = np;
    }
    uint64_t* kb[2] = {s->sort_keys, s->sort_keys + np};
    uint32_t* rb[2] = {s->sort_rows, s->sort_rows + np};
    uint32_t* totals = s->sort_hist + (size_t)256 * tiles;
    pdl(k_sort_init, grid_for(s, np, 256), 256, 0, s)(s->m, np, kb[0], rb[0], masked ? s->mask : (const uint8_t*)nullptr); s->launches++;
    for (int pass = 0; pass < 8; ++pass) {
        const int a = pass & 1, b = a ^ 1;
        pdl(k_sort_hist, tiles, kSortThreads, 0, s)(kb[a], np, 8 * pass, s->sort_hist);
        pdl(k_sort_scan, 256, 256, 0, s)(s->sort_hist, tiles, totals);
        pdl(k_sort_scatter, tiles, kSortThreads, 0, s)(kb[a], rb[a], np, 8 * pass, s->sort_hist, totals, kb[b], rb[b]);
        s->launches += 3;
    }
    CU(cudaGetLastError());
    return ITAL_OK;
}

int64_t ital_top_results(ital_shard* s, int64_t k, int64_t* out_idx, double* out_val) {
    if (!s || !out_idx) return fail(ITAL_EINVAL, "ital_top_results: bad arguments");
    if (s->W == 0) return fail(ITAL_ESTATE, "ital_top_results before any labelled point");
    CU(cudaSetDevice(s->device));
    // pool rows only: query rows (global index >= n_data) are the last local rows of the last shard
    const int64_t np = std::max<int64_t>(0, std::min<int64_t>(s->n, s->n_data - s->row_offset));
    if (k < 0 || k > np) k = np;
    if (k == 0) return 0;
    int rc = sort_rows_by_mean(s, np, false);
    if (rc) return rc;
    pdl(k_sort_gather, grid_for(s, k, 256), 256, 0, s)(s->sort_rows, k, s->row_offset, s->m, s->sort_out_idx,
                                                       s->sort_out_val); s->launches++;
    CU(cudaGetLastError());
    CU(copy_async(s, out_idx, s->sort_out_idx, (size_t)k * sizeof(int64_t), cudaMemcpyDeviceToHost, s->stream));
    if (out_val)
        CU(copy_async(s, out_val, s->sort_out_val, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return k;
}

// ITAL.fetch_unlabelled's top_candidates restriction (ital/ital.py:111-117) on the device: the `top` unseen local pool
// rows with the largest posterior mean stay candidates (exact ties at the cut go to the lower row), everything else
// gets the restricted bit.  Nothing is copied to the host.
int ital_restrict_top(ital_shard* s, int64_t top) {
    if (!s || top < 1) return fail(ITAL_EINVAL, "ital_restrict_top: bad arguments");
    if (s->W == 0) return fail(ITAL_ESTATE, "ital_restrict_top before any labelled point");
    CU(cudaSetDevice(s->device));
    const int64_t np = std::max<int64_t>(0, std::min<int64_t>(s->n, s->n_data - s->row_offset));
    if (np == 0) return ITAL_OK;
    int rc = sort_rows_by_mean(s, np, true);            // (masks without the restricted bit: lifted before)
    if (rc) return rc;
    pdl(k_mask_all, grid_for(s, s->n, 256), 256, 0, s)(s->mask, s->n, 0xff, kRestricted); s->launches++;
    pdl(k_mask_clear_sorted, grid_for(s, std::min(top, np), 256), 256, 0, s)(s->mask, s->sort_rows, std::min(top, np),
                                                                            kRestricted); s->launches++;
    CU(cudaGetLastError());
    return ITAL_OK;
}

static int predict_impl(ital_shard* s, const double* Xt, int64_t mrows, double* out_mean, double* out_var, double* out_proj) {
    if (!s || !Xt || mrows < 0 || !out_mean) return fail(ITAL_EINVAL, "ital_predict: bad arguments");
    if (s->W == 0) return fail(ITAL_ESTATE, "ital_predict before any labelled point");
    if (mrows == 0) return ITAL_OK;
    CU(cudaSetDevice(s->device));
    const int nl = s->W;
    if (!s->w_valid) {              // w = K^-1 y = L^-T (L^-1 y)  (gp.py:158,196), from the device-resident factor
        const size_t wsm = (size_t)nl * sizeof(double);
        if (wsm > 48 * 1024) CU(cudaFuncSetAttribute(k_model_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
        pdl(k_model_w, 1, 1024, wsm, s)(model_refs(s), nl, s->w_vec_dev); s->launches++;
        CU(cudaGetLastError());
        s->w_valid = true;
    }
    double *xt_dev = nullptr, *mean_dev = nullptr, *var_dev = nullptr, *proj_dev = nullptr;
    CU(cudaMalloc(&xt_dev, (size_t)mrows * s->d * sizeof(double)));
    CU(cudaMalloc(&mean_dev, (size_t)mrows * sizeof(double)));
    if (out_var) CU(cudaMalloc(&var_dev, (size_t)mrows * sizeof(double)));
    if (out_proj) CU(cudaMalloc(&proj_dev, (size_t)mrows * nl * sizeof(double)));
    CU(copy_async(s, xt_dev, Xt, (size_t)mrows * s->d * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    const int threads = 128, wpb = threads / 32;
    const size_t smem = (size_t)wpb * nl * sizeof(double);
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pdl(k_predict, (unsigned)((mrows + wpb - 1) / wpb), threads, smem, s)(
        xt_dev, mrows, (int)s->d, s->lab_x_dev, s->lab_sqn_dev, nl, s->w_vec_dev, s->LK_dev, (int64_t)s->model_cap, s->var,
        -2.0 * s->ls * s->ls, mean_dev, var_dev, proj_dev); s->launches++;
    CU(cudaGetLastError());
    CU(copy_async(s, out_mean, mean_dev, (size_t)mrows * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out_var) CU(copy_async(s, out_var, var_dev, (size_t)mrows * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out_proj) CU(copy_async(s, out_proj, proj_dev, (size_t)mrows * nl * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaFree(xt_dev));
    CU(cudaFree(mean_dev));
    if (var_dev) CU(cudaFree(var_dev));
    if (proj_dev) CU(cudaFree(proj_dev));
    return ITAL_OK;
}

int ital_predict(ital_shard* s, const double* Xt, int64_t mrows, double* out_mean, double* out_var) {
    return predict_impl(s, Xt, mrows, out_mean, out_var, nullptr);
}

int ital_predict_proj(ital_shard* s, const double* Xt, int64_t mrows, double* out_mean, double* out_proj) {
    if (!out_proj) return fail(ITAL_EINVAL, "ital_predict_proj: bad arguments");
    return predict_impl(s, Xt, mrows, out_mean, nullptr, out_proj);
}

int ital_set_lazy_rows(ital_shard* s, int on) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    if (s->fetching) return fail(ITAL_ESTATE, "ital_set_lazy_rows during a fetch");
    s->lazy_rows = on != 0;
    return ITAL_OK;
}

int ital_set_label_estimation(ital_shard* s, int mode) {
    if (!s || mode < 0 || mode > 2) return fail(ITAL_EINVAL, "ital_set_label_estimation: mode must be 0, 1 or 2");
    if (s->fetching) return fail(ITAL_ESTATE, "ital_set_label_estimation during a fetch");
    s->estimation = mode;
    return ITAL_OK;
}

int ital_set_fused(ital_shard* s, int on) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    if (s->fetching) return fail(ITAL_ESTATE, "ital_set_fused during a fetch");
    s->fused = on != 0;
    return ITAL_OK;
}

int64_t ital_fused_trace(ital_shard* s, int on, uint64_t* out, int64_t max_out) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    CU(cudaSetDevice(s->device));
    constexpr int kMarks = 64;
    if (!s->f_trace) {
        CU(cudaMalloc(&s->f_trace, kMarks * sizeof(unsigned long long)));
        CU(cudaMemset(s->f_trace, 0, kMarks * sizeof(unsigned long long)));
    }
    s->f_trace_on = on != 0;
    if (!out) return 0;
    CU(cudaStreamSynchronize(s->stream));
    const int64_t m = std::min<int64_t>(max_out, kMarks);
    CU(cudaMemcpy(out, s->f_trace, (size_t)m * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return m;
}

int ital_set_bulk_stream(ital_shard* s, int on) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    s->bulk_stream = on != 0;
    return ITAL_OK;
}

int ital_profile_enable(ital_shard* s, int on) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    s->profiling = on != 0;
    return ITAL_OK;
}

int ital_profile_read(ital_shard* s, double* ms_total, int64_t* launches, double* algorithmic_bytes) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    double total = 0.0;
    for (auto& pr : s->prof_events) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (ms_total) *ms_total = total;
    if (launches) *launches = (int64_t)s->prof_events.size();
    if (algorithmic_bytes) *algorithmic_bytes = s->prof_bytes;
    s->prof_events.clear();
    s->prof_bytes = 0.0;
    return ITAL_OK;
}

int64_t ital_launch_count(const ital_shard* s) { return s ? s->launches : 0; }

int ital_transfer_bytes(const ital_shard* s, int64_t* h2d, int64_t* d2h) {
    if (!s) return fail(ITAL_EINVAL, "null shard");
    if (h2d) *h2d = s->h2d_bytes;
    if (d2h) *d2h = s->d2h_bytes;
    return ITAL_OK;
}

int64_t ital_snq_nodes(int t, const double* m, const double* L, double* eta, double* w, int32_t* orth, double* masses) {
    if (t < 1 || t > 10 || !m || !L) return fail(ITAL_EINVAL, "ital_snq_nodes: bad arguments");
    if (!eta) return snq::capacity_for(t);
    snq::Nodes nd = snq::generate(t, m, L);
    memcpy(eta, nd.eta.data(), nd.eta.size() * sizeof(double));
    if (w) memcpy(w, nd.w.data(), nd.w.size() * sizeof(double));
    if (orth) memcpy(orth, nd.orth.data(), nd.orth.size() * sizeof(int32_t));
    if (masses) memcpy(masses, nd.masses.data(), nd.masses.size() * sizeof(double));
    return nd.n;
}

int ital_snq_order(int t) { return snq::order_for(t); }

int64_t ital_phi_table(double* out, int64_t max_out) {
    const std::vector<double> tab = build_phi_table();
    if (out) memcpy(out, tab.data(), (size_t)std::min<int64_t>(max_out, (int64_t)tab.size()) * sizeof(double));
    return (int64_t)tab.size();
}

int64_t ital_h_table(double* out, int64_t max_out) {
    const std::vector<double> tab = build_h_table();
    if (out) memcpy(out, tab.data(), (size_t)std::min<int64_t>(max_out, (int64_t)tab.size()) * sizeof(double));
    return (int64_t)tab.size();
}

int ital_snq_sub(int n_batch, int n_cols, const double* m, const double* L, double noise, int64_t* sizes, double* eta,
                 double* w, int32_t* group_begin, double* tables) {
    if (n_batch < 0 || n_batch > 7 || n_cols < n_batch || n_cols > kSubMaxCols || n_cols - n_batch > kSubMaxU ||
        !m || !L || !sizes)
        return fail(ITAL_EINVAL, "ital_snq_sub: bad arguments");
    snq::SubSets ss = snq::generate_sub(n_batch, n_cols, m, L, noise);
    const int G = 1 << n_batch, D = n_cols, u = n_cols - n_batch;
    sizes[0] = ss.n_nodes;
    sizes[1] = ss.n_groups;
    sizes[2] = ss.sub_bits;
    sizes[3] = 2 * G + (int64_t)G * D + D * D + (int64_t)G * u + u * u + u * D;
    if (!eta) return ITAL_OK;
    memcpy(eta, ss.eta.data(), ss.eta.size() * sizeof(double));
    memcpy(w, ss.w.data(), ss.w.size() * sizeof(double));
    memcpy(group_begin, ss.group_begin.data(), ss.group_begin.size() * sizeof(int32_t));
    double* p = tables;
    auto put = [&](const std::vector<double>& v, size_t count) {
        if (count) memcpy(p, v.data(), count * sizeof(double));
        p += count;
    };
    put(ss.mass, 2 * G);
    put(ss.mu, (size_t)G * D);
    put(ss.Sig, (size_t)D * D);
    put(ss.mU, (size_t)G * u);
    put(ss.CU, (size_t)u * u);
    put(ss.BS, (size_t)u * D);
    return ITAL_OK;
}

int ital_snq_general(int t, const double* m, const double* L, double noise, int64_t* sizes, double* eta, double* w,
                     int32_t* group_begin, double* group_mass, int32_t* set_group0, int32_t* lut) {
    if (t < 1 || t > 4 || !m || !L || !sizes) return fail(ITAL_EINVAL, "ital_snq_general: bad arguments");
    snq::GeneralSets gs = snq::generate_general(t, m, L, noise);
    sizes[0] = gs.n_nodes;
    sizes[1] = gs.n_groups;
    sizes[2] = gs.n_sets;
    sizes[3] = (int64_t)gs.lut.size();
    if (!eta) return ITAL_OK;
    memcpy(eta, gs.eta.data(), gs.eta.size() * sizeof(double));
    memcpy(w, gs.w.data(), gs.w.size() * sizeof(double));
    memcpy(group_begin, gs.group_begin.data(), gs.group_begin.size() * sizeof(int32_t));
    memcpy(group_mass, gs.group_mass.data(), gs.group_mass.size() * sizeof(double));
    memcpy(set_group0, gs.set_group0.data(), gs.set_group0.size() * sizeof(int32_t));
    memcpy(lut, gs.lut.data(), gs.lut.size() * sizeof(int32_t));
    return ITAL_OK;
}

}  // extern "C"
