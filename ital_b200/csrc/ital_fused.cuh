// The whole greedy loop of ITAL.fetch_unlabelled (ital/ital.py:119-134) as ONE persistent cooperative kernel.
//
// Applies when the lazy-greedy bound prunes (users who label every sample, label_prob >= 1; not `exhaustive`) and
// the batch-conditional projections are extended on demand (lazy rows): then a greedy step touches a few hundred rows
// of the pool, the n-sized work is two scans of (mask, gain), and what is left of a multi-kernel fetch is launch
// latency between ~30 small dependent kernels.  Here one CTA per SM stays resident for the whole fetch and the phases
// of a step are separated by grid-wide barriers:
//
//   S0   closed-form scores of the first step for every candidate (k_score0), gain = score, argmax, commit
//   per step t = 1 .. k-1 (k <= 4):
//   P1   quadrature nodes of the step, one chunk per CTA (snq_node);  stage A: every 256-thread team takes the maximum
//        of the bound inside its strided subset of the pool and brings that row's projection up to date (catchup_row)
//   P2   order-preserving compaction of the kept nodes into the global node arrays
//   P3   base orthant masses / H(base) (every CTA, same order);  exact score of the team's stage-A row (eval_candidate)
//   P4   threshold from the best stage-A score, worklist of the rows whose bound still reaches it (k_worklist)
//   P5   the worklist rows: projections caught up by one warp per row, eight rows of a team in flight, then scored
//   P6   argmax over the teams' bests, the winner's record into every CTA's shared memory, batch state committed
//
// The arithmetic is that of the multi-kernel path, function by function (score0_value, snq_node, chunk_mass_warp + masses_from_chunks,
// catchup_row, eval_candidate with 256-thread teams), so both paths return the same batch with bit-identical scores.
// Multi-GPU: in P6 CTA 0 stores the shard's proposal into its peers' exchange buffers (peer_put, NVLink) and every
// CTA waits for the peers' flags in local memory -- the exchange of ital_fetch_peer without its two extra launches.
#pragma once
#include "ital_kernels.cuh"

namespace italk {

constexpr int kFusedThreads = 512;
constexpr int kFusedTeam = 256;
constexpr int kFusedTeams = kFusedThreads / kFusedTeam;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedMaxSteps = 4;               // greedy steps with at most 3 base variables (tensor rule on the device)

struct FusedArgs {
    // pool
    const void* X;
    int64_t n;
    int d, d_pad;
    int64_t row_offset;
    const double* sqn;
    const double* m;
    const double* v;
    double* U;
    int64_t ldu;
    const uint8_t* mask;
    uint8_t* mask_rw;               // != nullptr: mark the selected rows (a multi-kernel continuation follows)
    uint8_t sel_bits;
    double* gain;
    double* score;
    uint32_t* tags;
    uint32_t epoch;
    int W, w_cap, k;
    double var, neg2ls2, log1p_eps, flag_var, margin;
    double shift_coef[kFusedMaxSteps];  // mistaken user: additive constant of step t per unit of total mass
    // quadrature
    const double* gl_x;
    const double* gl_w;
    const double2* phi;
    const double* htab;             // table of h_tab (closed-form first step)
    double R, w_min;
    int q_min;
    int order[kFusedMaxSteps];
    int64_t node_cap[kFusedMaxSteps];
    int chunk_cap;                  // nodes generated per CTA at most (shared-memory staging)
    double* nodes4;                 // {eta_0, eta_1, eta_2, weight} per kept node, zero-padded to kNodePad
    int* group_begin;               // [9] orthant offsets (for a multi-kernel continuation)
    double* masses;
    double* hbase;
    // batch state in global memory (written by CTA 0; read by the host and by a multi-kernel continuation)
    double* rec_hist;
    int64_t rec_len;
    double* rec_in;
    double* base_m;
    double* base_L;
    double* sel;
    int* stats;
    int* counters;
    // scratch
    Best* blk_best;                 // [gridDim.x * kFusedTeams]
    int* blk_cnt;                   // [gridDim.x][8] kept nodes per orthant in every CTA's chunk
    double* blk_mass;               // [gridDim.x][8] orthant masses of every CTA's chunk of nodes
    int* stage_rows;                // [gridDim.x * kFusedTeams]
    int* worklist;
    unsigned* barrier;
    unsigned bar_base;
    int want_scores;
    // peer exchange (multi-GPU)
    PeerPut pp;                     // pp.epoch = epoch of the first step; step t uses pp.epoch + t
    const unsigned long long* flags;
    const double* slots;            // base of the local slots: [2][G][slot_doubles]
    int* peer_error;
    unsigned long long* trace;      // != nullptr: CTA 0 stamps %globaltimer at every phase boundary (diagnostics)
};

// shared-memory carve-up of k_fetch_fused, in doubles (host and device use the same numbers)
struct FusedSmem {
    size_t recs, uv, red, phi, part, masses, hb, base_m, base_L, ss, si, nd_eta, nd_w, nd_orth, ism, sel_loc, total;
    __host__ __device__ FusedSmem(int64_t rec_len, int w_cap, int C, int n_ctas) {
        size_t o = 0;
        recs = o; o += (size_t)kFusedMaxSteps * rec_len;
        uv = o; o += (size_t)kFusedWarps * w_cap;
        red = o; o += kFusedTeams * 64;
        phi = o; o += 2 * (size_t)kPhiTableLen;     // table of phi_tab
        part = o; o += (size_t)n_ctas * 8;          // chunk sums of the base masses
        masses = o; o += 8;
        hb = o; o += 2;
        base_m = o; o += kFusedMaxSteps;
        base_L = o; o += kFusedMaxSteps * kBaseStride;
        ss = o; o += kFusedTeams * 16;
        si = o; o += kFusedTeams * 16;
        nd_eta = o; o += 3 * (size_t)C;
        nd_w = o; o += C;
        nd_orth = o; o += (C + 1) / 2;
        ism = o; o += 128;                      // 256 ints (the first 16 also serve as orthant offsets)
        sel_loc = o; o += kFusedMaxSteps;
        total = o;
    }
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid-wide barrier of a cooperative launch: one arrival per CTA on a monotone counter.  Release/acquire at GPU
// scope around the CTA barriers orders every global write before it against every read after it.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while ((int)(ld_acquire_gpu_u32(bar) - target) < 0) {
        }
        __threadfence();
    }
    __syncthreads();
}

// argmax over a group of threads (a team on a named barrier, or the whole CTA on barrier 0); every thread returns it
__device__ __noinline__ Best group_argmax(double bs, long long bi, int tid_group, int nthreads, int bar_id,
                                             double* ss, long long* si) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(os, oi, bs, bi)) { bs = os; bi = oi; }
    }
    if ((tid_group & 31) == 0) { ss[tid_group >> 5] = bs; si[tid_group >> 5] = bi; }
    team_barrier(bar_id, nthreads);
    Best r;
    r.score = ss[0];
    r.idx = si[0];
    for (int w = 1; w < nthreads / 32; ++w)
        if (better(ss[w], si[w], r.score, r.idx)) { r.score = ss[w]; r.idx = si[w]; }
    team_barrier(bar_id, nthreads);
    return r;
}

// One copy of the row-level routines inside the persistent kernel (inlined at every call site they made half a
// megabyte of SASS, and a kernel that runs every piece of its code once per launch is bound by instruction fetch).
template <int T>
__device__ __noinline__ void eval_candidate_call(const EvalArgs& a, int64_t i, int tid_team, int TPC, int bar_id,
                                                 double* red, const double2* phi, const int* gb, const double* masses,
                                                 double h_base) {
    eval_candidate<T>(a, i, tid_team, TPC, bar_id, red, phi, gb, masses, h_base);
}

__device__ __forceinline__ void eval_dispatch(int tn, const EvalArgs& a, int64_t i, int tid_team, int TPC, int bar_id,
                                              double* red, const double2* phi, const int* gb, const double* masses,
                                              double h_base) {
    if (tn == 1) eval_candidate_call<1>(a, i, tid_team, TPC, bar_id, red, phi, gb, masses, h_base);
    else if (tn == 2) eval_candidate_call<2>(a, i, tid_team, TPC, bar_id, red, phi, gb, masses, h_base);
    else eval_candidate_call<3>(a, i, tid_team, TPC, bar_id, red, phi, gb, masses, h_base);
}

template <typename XT>
__device__ __noinline__ void catchup_row_call(int64_t i, int lane, const XT* X, int d, int d_pad, const double* recs,
                                              int64_t rec_len, int w_cap, int W, int t, const double* sqn, double* U,
                                              int64_t ldu, uint32_t* tags, uint32_t epoch, double var, double neg2ls2,
                                              double* uv) {
    catchup_row<XT>(i, lane, X, d, d_pad, recs, rec_len, w_cap, W, t, sqn, U, ldu, tags, epoch, var, neg2ls2, uv);
}

template <int T>
__device__ __noinline__ void snq_node_call(int64_t k, int64_t N, int q, double R, int q_min, const double* base_m,
                                           const double* base_L, const double* gl_x, const double* gl_w, double* e,
                                           double& wt, int& ob) {
    snq_node<T>(k, N, q, R, q_min, base_m, base_L, gl_x, gl_w, e, wt, ob);
}

template <typename XT>
__global__ void __launch_bounds__(kFusedThreads, 1) k_fetch_fused(FusedArgs a) {
    extern __shared__ __align__(16) double fsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int team = tid / kFusedTeam, tid_team = tid % kFusedTeam, warp_team = tid_team >> 5;
    const int n_teams = gridDim.x * kFusedTeams, gt = blockIdx.x * kFusedTeams + team;
    const int team_bar = 1 + team;
    const XT* X = (const XT*)a.X;
    const int C = a.chunk_cap;
    unsigned target = a.bar_base;
    int n_marks = 0;
#define FUSED_MARK()                                                                     \
    do {                                                                                 \
        if (a.trace != nullptr && blockIdx.x == 0 && tid == 0) a.trace[n_marks] = global_ns(); \
        ++n_marks;                                                                       \
    } while (0)
    FUSED_MARK();

    // shared memory
    const FusedSmem L(a.rec_len, a.w_cap, C, (int)gridDim.x);
    double* recs = fsm + L.recs;                                        // [kFusedMaxSteps][rec_len]
    double* uv = fsm + L.uv + (size_t)warp * a.w_cap;                   // [warps][w_cap]
    double* red = fsm + L.red + team * 64;                              // [teams][64]
    double2* phi_s = reinterpret_cast<double2*>(fsm + L.phi);           // table of phi_tab
    double* chunk_part = fsm + L.part;                                  // [gridDim.x][8]
    double* masses = fsm + L.masses;                                    // [8]
    double* hb = fsm + L.hb;                                            // [2]
    double* base_m = fsm + L.base_m;                                    // [4]
    double* base_L = fsm + L.base_L;                                    // [4][kBaseStride]
    double* ss = fsm + L.ss;                                            // [teams][16]
    long long* si = reinterpret_cast<long long*>(fsm + L.si);           // [teams][16]
    double* nd_eta = fsm + L.nd_eta;                                    // [3][C]
    double* nd_w = fsm + L.nd_w;                                        // [C]
    int* nd_orth = reinterpret_cast<int*>(fsm + L.nd_orth);             // [C]
    int* ism = reinterpret_cast<int*>(fsm + L.ism);                     // [256] small integers
    int* gbeg = ism + 16;                                               // [9] orthant offsets of the step's nodes
    long long* sel_loc = reinterpret_cast<long long*>(fsm + L.sel_loc); // [4] local rows selected so far (-1: remote)

    if (blockIdx.x == 0 && tid < 4) {
        a.counters[tid] = 0;
        a.stats[tid] = 0;
    }
    if (tid < 2) hb[tid] = (double)tid;                                  // first step: no base, total mass 1
    if (tid < kFusedMaxSteps) sel_loc[tid] = -1;
    phi_tab_to_shared(phi_s, a.phi);

    // ---- S0: closed form for every candidate ------------------------------------------------------------------
    Best win;
    {
        double bs = 0.0;
        long long bi = -1;
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + tid; i0 < a.n; i0 += 2 * stride) {
            uint8_t mk[2];
            double mm[2], vv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t i = i0 + u * stride;
                mk[u] = i < a.n ? a.mask[i] : (uint8_t)1;
                mm[u] = i < a.n ? a.m[i] : 0.0;
                vv[u] = i < a.n ? a.v[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t i = i0 + u * stride;
                if (i >= a.n) break;
                double s = nan("");
                if (mk[u] == 0) {
                    s = score0_value(mm[u], vv[u], a.log1p_eps, 1.0, a.htab);
                    a.gain[i] = s;
                    if (better(s, i, bs, bi)) { bs = s; bi = i; }
                }
                if (a.want_scores) a.score[i] = s;
            }
        }
        const Best b = group_argmax(bs, bi, tid, kFusedThreads, 0, ss, si);
        if (tid == 0) a.blk_best[blockIdx.x] = b;
        if (warp == 0 && b.idx >= 0)
            prefetch_row<XT>(b.idx, lane, X, a.d_pad, a.U, a.ldu, a.W, a.m, a.v, a.sqn, a.gain, a.tags);
    }
    FUSED_MARK();
    grid_barrier(a.barrier, target);
    FUSED_MARK();
    {
        double bs = 0.0;
        long long bi = -1;
        for (int k = tid; k < (int)gridDim.x; k += blockDim.x) {
            const double s = __ldcg(&a.blk_best[k].score);
            const long long i = __ldcg(&a.blk_best[k].idx);
            if (better(s, i, bs, bi)) { bs = s; bi = i; }
        }
        win = group_argmax(bs, bi, tid, kFusedThreads, 0, ss, si);
    }

    // stage A of a step: every team takes the maximum of the bound (gain) in its strided subset of the pool, leaving
    // out the rows in excl[] (selected so far)
    auto stage_a_scan = [&](const int* excl) -> Best {
        double bs = 0.0;
        long long bi = -1;
        const int64_t stride = (int64_t)n_teams * kFusedTeam;
        for (int64_t i0 = (int64_t)gt * kFusedTeam + tid_team; i0 < a.n; i0 += 4 * stride) {
            uint8_t mk[4];
            double val[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                mk[u] = i < a.n ? a.mask[i] : (uint8_t)1;
                val[u] = i < a.n ? __ldcg(a.gain + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                const int ii = (int)i;
                const bool ok = mk[u] == 0 && ii != excl[0] && ii != excl[1] && ii != excl[2] && ii != excl[3];
                // (score desc, row asc; a NaN bound never wins) -- rows come in ascending order per thread
                if (ok && (bi < 0 ? val[u] == val[u] : val[u] > bs)) { bs = val[u]; bi = i; }
            }
        }
        return group_argmax(bs, bi, tid_team, kFusedTeam, team_bar, ss + team * 16, si + team * 16);
    };
    bool dead = false;
    for (int t = 0; t < a.k; ++t) {
        bool have_early = false;
        Best early_b;
        early_b.score = 0.0;
        early_b.idx = -1;
        long long early_prop = -1;
        // ---- commit the winner of step t (np.argmax + AppendedMutualInformation.append, ital.py:130-131) ----------
        // every CTA builds the winner's record in its own shared memory; CTA 0 also writes the global batch state
        double* rec = recs + (size_t)t * a.rec_len;
        long long row = dead ? -1 : win.idx;
        double score = 0.0;
        if (row >= 0) {
            score = win.score + a.shift_coef[t] * hb[1];
            if (score != score) row = -1;                               // a NaN score never wins
        }
        const int Wt = a.W + t;
        {
            if (row >= 0) {
                for (int j = tid; j < a.w_cap; j += blockDim.x)
                    rec[8 + j] = j < Wt ? __ldcg(a.U + (int64_t)j * a.ldu + row) : 0.0;
                for (int j = tid; j < a.d; j += blockDim.x) rec[8 + a.w_cap + j] = (double)X[row * (int64_t)a.d_pad + j];
                if (tid == 0) {
                    double cv = a.v[row];
                    for (int j = a.W; j < Wt; ++j) {
                        const double e = __ldcg(a.U + (int64_t)j * a.ldu + row);
                        cv = fma(-e, e, cv);
                    }
                    rec[0] = (double)(a.row_offset + row);
                    rec[1] = score;
                    rec[2] = a.m[row];
                    rec[3] = cv;
                    rec[4] = a.sqn[row];
                    rec[5] = a.v[row];
                    rec[6] = __ldcg(a.gain + row);
                    rec[7] = 0.0;
                }
            } else {
                for (int j = tid; j < (int)a.rec_len; j += blockDim.x) rec[j] = j == 0 ? -1.0 : (j == 1 ? -INFINITY : 0.0);
            }
        }
        __syncthreads();
        long long local_row = row;                                      // the winner as a local row, or -1
        if (a.pp.peer_base != nullptr) {
            // exchange: this shard's proposal into the slot reserved for it in every shard's buffer, then wait for all
            PeerPut pp = a.pp;
            pp.epoch = a.pp.epoch + (unsigned long long)t;
            if (blockIdx.x == 0) peer_put(pp, rec, a.rec_len);
            if (t + 1 < a.k && row >= 0) {
                // while the proposals travel: stage A of the next step needs the gains only, not the winner
                int excl[kFusedMaxSteps];
#pragma unroll
                for (int c = 0; c < kFusedMaxSteps; ++c) excl[c] = c < t ? (int)sel_loc[c] : -1;
                excl[t] = (int)row;                                     // (t < kFusedMaxSteps - 1 here)
                // CTA 0 is busy with the exchange itself (record, stores, system fence): its two teams sit this
                // stage A out -- the sample only sets the pruning threshold, 2 of 296 subsets less do not matter
                if (blockIdx.x != 0) early_b = stage_a_scan(excl);
                early_prop = row;
                have_early = true;
            }
            if (tid < pp.G) {
                const unsigned long long t0 = global_ns();
                unsigned spins = 0;
                while (ld_acquire_sys(a.flags + tid) < pp.epoch) {
                    if ((++spins & 1023u) == 0 && global_ns() - t0 > 5000000000ull) { *a.peer_error = 1; break; }
                }
            }
            __syncthreads();
            const double* slots = a.slots + (int64_t)(pp.epoch & 1) * pp.G * pp.slot_doubles;
            if (tid == 0) {
                int wg = -1;
                for (int g = 0; g < pp.G; ++g) {
                    const double idx = __ldcg(slots + g * pp.slot_doubles), sc = __ldcg(slots + g * pp.slot_doubles + 1);
                    if (idx < 0.0 || sc != sc) continue;
                    if (wg < 0 || sc > __ldcg(slots + wg * pp.slot_doubles + 1) ||
                        (sc == __ldcg(slots + wg * pp.slot_doubles + 1) && idx < __ldcg(slots + wg * pp.slot_doubles)))
                        wg = g;
                }
                ism[32] = wg;
            }
            __syncthreads();
            const int wg = ism[32];
            if (wg < 0) {
                for (int j = tid; j < (int)a.rec_len; j += blockDim.x) rec[j] = j == 0 ? -1.0 : (j == 1 ? -INFINITY : 0.0);
                local_row = -1;
                row = -1;
            } else {
                const double* r = slots + (int64_t)wg * pp.slot_doubles;
                for (int j = tid; j < (int)a.rec_len; j += blockDim.x) rec[j] = __ldcg(r + j);
                const long long loc = (long long)__ldcg(r) - a.row_offset;
                local_row = (loc >= 0 && loc < a.n) ? loc : -1;
                row = 0;                                                // some shard has a winner
            }
            __syncthreads();
        }
        if (row < 0) dead = true;
        if (tid == 0) {
            sel_loc[t] = dead ? -1 : local_row;
            base_m[t] = rec[2];
            for (int j = 0; j < t; ++j) base_L[t * kBaseStride + j] = rec[8 + a.W + j];
            base_L[t * kBaseStride + t] = sqrt(fmax(rec[3], 1e-300));
        }
        if (blockIdx.x == 0) {
            double* hist = a.rec_hist + (int64_t)t * a.rec_len;
            for (int j = tid; j < (int)a.rec_len; j += blockDim.x) {
                hist[j] = rec[j];
                a.rec_in[j] = rec[j];
            }
            if (tid == 0) {
                a.sel[2 * t] = rec[0];
                a.sel[2 * t + 1] = rec[1];
                if (!dead) {
                    a.base_m[t] = rec[2];
                    for (int j = 0; j < t; ++j) a.base_L[t * kBaseStride + j] = rec[8 + a.W + j];
                    a.base_L[t * kBaseStride + t] = sqrt(fmax(rec[3], 1e-300));
                    if (a.mask_rw != nullptr && local_row >= 0) a.mask_rw[local_row] |= a.sel_bits;
                }
            }
        }
        __syncthreads();
        FUSED_MARK();
        if (t + 1 >= a.k) break;
        const int tn = t + 1;                                           // the step scored next: tn base variables
        const int64_t N = a.node_cap[tn];

        // ---- P1: nodes of the step (one chunk per CTA) and stage A -------------------------------------------------
        bool my_keep = false;
        int my_rank = 0, my_ob = 0;
        if (!dead) {
            const int64_t per = (N + gridDim.x - 1) / gridDim.x;        // <= C
            const int64_t k = (int64_t)blockIdx.x * per + tid;
            const bool have = tid < per && k < N;
            double e[3] = {0.0, 0.0, 0.0};
            double wt = 0.0;
            int ob = 0;
            if (have) {
                if (tn == 1) snq_node_call<1>(k, N, a.order[1], a.R, a.q_min, base_m, base_L, a.gl_x, a.gl_w, e, wt, ob);
                else if (tn == 2) snq_node_call<2>(k, N, a.order[2], a.R, a.q_min, base_m, base_L, a.gl_x, a.gl_w, e, wt, ob);
                else snq_node_call<3>(k, N, a.order[3], a.R, a.q_min, base_m, base_L, a.gl_x, a.gl_w, e, wt, ob);
            }
            const bool keep = have && wt >= a.w_min;
            // staged in generation order; rank of every kept node among the kept nodes of its orthant in this chunk
            if (tid < C) {
                nd_eta[tid] = e[0];
                nd_eta[C + tid] = e[1];
                nd_eta[2 * C + tid] = e[2];
                nd_w[tid] = keep ? wt : 0.0;
                nd_orth[tid] = ob;
            }
            int rank_w = 0;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const unsigned bal = __ballot_sync(0xffffffffu, keep && ob == o);
                if (ob == o) rank_w = __popc(bal & ((1u << lane) - 1u));
                if (lane == 0) ism[64 + warp * 8 + o] = __popc(bal);
            }
            __syncthreads();
            if (keep) {
                int r = rank_w;
                for (int ww = 0; ww < warp; ++ww) r += ism[64 + ww * 8 + ob];
                my_rank = r;
            }
            if (tid < 8) {
                int c = 0;
                for (int ww = 0; ww < kFusedWarps; ++ww) c += ism[64 + ww * 8 + tid];
                a.blk_cnt[blockIdx.x * 8 + tid] = c;
            }
            if (warp == 1) chunk_mass_warp(nd_w, nd_orth, (int)min((int64_t)per, max((int64_t)0, N - (int64_t)blockIdx.x * per)),
                                           lane, 0.0, a.blk_mass + (size_t)blockIdx.x * 8);
            my_keep = keep;
            my_ob = ob;
        }
        int selr[kFusedMaxSteps];                                       // rows selected so far never compete again
#pragma unroll
        for (int c = 0; c < kFusedMaxSteps; ++c) selr[c] = c <= t ? (int)sel_loc[c] : -1;
        long long a_row = -1;                                           // the team's stage-A row
        if (!dead) {
            Best b;
            if (have_early) {
                // the scan ran while the proposals were on their way (in the exchange above, several GPUs): it left out this shard's own
                // proposal, which is still a candidate if another shard's won -- the team that owns it adds it back
                b = early_b;
                int owner = (int)((early_prop / kFusedTeam) % n_teams);
                if (owner < kFusedTeams && n_teams > 2 * kFusedTeams) owner += kFusedTeams;     // (not a team of CTA 0)
                if (early_prop >= 0 && early_prop != sel_loc[t] && owner == gt) {
                    const double gv = __ldcg(a.gain + early_prop);
                    if (gv == gv && (b.idx < 0 || better(gv, early_prop, b.score, b.idx))) { b.score = gv; b.idx = early_prop; }
                }
            } else {
                b = stage_a_scan(selr);
            }
            a_row = b.idx;
            if (a_row >= 0 && warp_team == 1 && lane < 2) prefetch_l2(lane == 0 ? a.m + a_row : a.v + a_row);
            if (a_row >= 0 && warp_team == 0)
                catchup_row_call<XT>(a_row, lane, X, a.d, a.d_pad, recs, a.rec_len, a.w_cap, a.W, tn, a.sqn, a.U, a.ldu,
                                a.tags, a.epoch, a.var, a.neg2ls2, uv);
            if (tid_team == 0) a.stage_rows[gt] = (int)a_row;
        }
        FUSED_MARK();
        grid_barrier(a.barrier, target);
        FUSED_MARK();

        // ---- P2: compaction of the kept nodes (generation order) ---------------------------------------------------
        int NK = 0;
        if (!dead) {
            // per orthant: kept nodes in the chunks before this one, and in all chunks (warp o counts orthant o)
            if (warp < 8) {
                int before = 0, all = 0;
                for (int k = lane; k < (int)gridDim.x; k += 32) {
                    const int c = __ldcg(a.blk_cnt + k * 8 + warp);
                    all += c;
                    if (k < (int)blockIdx.x) before += c;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    before += __shfl_xor_sync(0xffffffffu, before, o);
                    all += __shfl_xor_sync(0xffffffffu, all, o);
                }
                if (lane == 0) { ism[32 + warp] = before; ism[40 + warp] = all; }
            }
            __syncthreads();
            if (tid == 0) {
                int g = 0, kept = 0;
                for (int o = 0; o < (1 << tn); ++o) {
                    gbeg[o] = g;
                    g += pad_nodes(ism[40 + o]);
                    kept += ism[40 + o];
                }
                gbeg[1 << tn] = g;
                ism[48] = kept;
            }
            __syncthreads();
            NK = ism[48];
            if (my_keep) {
                double* nd = a.nodes4 + 4 * (size_t)(gbeg[my_ob] + ism[32 + my_ob] + my_rank);
                nd[0] = nd_eta[tid];
                nd[1] = tn >= 2 ? nd_eta[C + tid] : 0.0;
                nd[2] = tn >= 3 ? nd_eta[2 * C + tid] : 0.0;
                nd[3] = nd_w[tid];
            }
            if ((int)blockIdx.x < (1 << tn)) {          // zero-weight tail of orthant blockIdx.x
                const int o = blockIdx.x;
                for (int k = gbeg[o] + ism[40 + o] + tid; k < gbeg[o + 1]; k += blockDim.x) {
                    double* nd = a.nodes4 + 4 * (size_t)k;
                    nd[0] = nd[1] = nd[2] = nd[3] = 0.0;
                }
            }
            if (blockIdx.x == 0 && tid <= (1 << tn)) a.group_begin[tid] = gbeg[tid];
        }
        FUSED_MARK();
        grid_barrier(a.barrier, target);
        FUSED_MARK();

        // ---- P3: base masses and H(base); exact score of the stage-A rows ------------------------------------------
        EvalArgs ea;
        ea.count = nullptr;
        ea.list = nullptr;
        ea.m = a.m;
        ea.v = a.v;
        ea.U = a.U;
        ea.ldu = a.ldu;
        ea.W0 = a.W;
        ea.eta = nullptr;
        ea.w = nullptr;
        ea.nodes4 = a.nodes4;
        ea.phi = a.phi;
        ea.orth = nullptr;
        ea.group_begin = nullptr;
        ea.n_nodes = N;
        ea.n_kept = nullptr;
        ea.masses = nullptr;
        ea.h_base = nullptr;
        ea.log1p_eps = a.log1p_eps;
        ea.flag_var = a.flag_var;
        ea.score = a.score;
        ea.gain = a.gain;
        ea.tags = a.tags;
        ea.epoch = a.epoch;
        ea.n_flagged = a.counters + 1;
        ea.n_scored = a.counters + 2;
        ea.force_block = 1;
        ea.t = tn;
        if (!dead) {
            for (int k = tid; k < (int)gridDim.x * 8; k += blockDim.x) chunk_part[k] = __ldcg(a.blk_mass + k);
            __syncthreads();
            masses_from_chunks(tn, (int)gridDim.x, chunk_part, a.log1p_eps, masses, hb);
            FUSED_MARK();
            if (blockIdx.x == 0 && tid < 8) {
                if (tid < (1 << tn)) a.masses[tid] = masses[tid];
                if (tid < 2) a.hbase[tid] = hb[tid];
                if (tid == 2) a.counters[3] = NK;
            }
            if (a_row >= 0 && tag_step(__ldcg(a.tags + a_row), a.epoch) != tn) {
                eval_dispatch(tn, ea, a_row, tid_team, kFusedTeam, team_bar, red, phi_s, gbeg, masses, hb[0]);
            }
        }
        FUSED_MARK();
        grid_barrier(a.barrier, target);
        FUSED_MARK();

        // ---- P4: threshold of the lazy-greedy bound, worklist ------------------------------------------------------
        if (!dead) {
            double bs = 0.0;
            long long bi = -1;
            for (int k = tid; k < n_teams; k += blockDim.x) {
                const long long i = __ldcg(a.stage_rows + k);
                if (i < 0) continue;
                const double s = __ldcg(a.score + i);
                if (better(s, i, bs, bi)) { bs = s; bi = i; }
            }
            const Best b = group_argmax(bs, bi, tid, kFusedThreads, 0, ss, si);
            double thr = -INFINITY;
            if (b.idx >= 0 && b.score == b.score) thr = b.score;
            thr = thr - a.margin - hb[0];
            const int64_t wstride = (int64_t)gridDim.x * blockDim.x;
            for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (tid & ~31); i0 < a.n; i0 += 2 * wstride) {
                uint8_t mk[2];
                double gv[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int64_t i = i0 + u * wstride + lane;
                    mk[u] = i < a.n ? a.mask[i] : (uint8_t)1;
                    gv[u] = i < a.n ? __ldcg(a.gain + i) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int64_t i = i0 + u * wstride + lane;
                    const int ii = (int)i;
                    const bool take = mk[u] == 0 && gv[u] >= thr && ii != selr[0] && ii != selr[1] && ii != selr[2] &&
                                      ii != selr[3];
                    unsigned ballot = __ballot_sync(0xffffffffu, take);
                    if (ballot != 0) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(a.counters, __popc(ballot));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (take) a.worklist[base + __popc(ballot & ((1u << lane) - 1u))] = ii;
                        while (ballot != 0) {                           // what P5 will read of these rows, into L2
                            const int src = __ffs(ballot) - 1;
                            ballot &= ballot - 1;
                            const int64_t r = __shfl_sync(0xffffffffu, ii, src);
                            prefetch_row<XT>(r, lane, X, a.d_pad, a.U, a.ldu, a.W + tn - 1, a.m, a.v, a.sqn, a.gain, a.tags);
                        }
                    }
                }
            }
        }
        FUSED_MARK();
        grid_barrier(a.barrier, target);
        FUSED_MARK();

        // ---- P5: exact scores of the worklist ----------------------------------------------------------------------
        {
            double bs = 0.0;
            long long bi = -1;
            if (!dead) {
                const int n_items = __ldcg(a.counters);
                if (n_items < 24 * (int)gridDim.x) {
                    // few rows: a 256-thread team per row; the projections of eight rows of a team are brought up to date
                    // by its eight warps at once (all their loads in flight), then the rows are scored one after the other
                    for (int r0 = 0; gt + (int64_t)r0 * n_teams < n_items; r0 += 8) {
                        const int64_t item_w = gt + (int64_t)(r0 + warp_team) * n_teams;
                        if (item_w < n_items) {
                            const int64_t i = __ldcg(a.worklist + item_w);
                            catchup_row_call<XT>(i, lane, X, a.d, a.d_pad, recs, a.rec_len, a.w_cap, a.W, tn, a.sqn, a.U,
                                            a.ldu, a.tags, a.epoch, a.var, a.neg2ls2, uv);
                        }
                        team_barrier(team_bar, kFusedTeam);
                        for (int rr = 0; rr < 8; ++rr) {
                            const int64_t item = gt + (int64_t)(r0 + rr) * n_teams;
                            if (item >= n_items) break;
                            const int64_t i = __ldcg(a.worklist + item);
                            if (tag_step(__ldcg(a.tags + i), a.epoch) != tn) {
                                eval_dispatch(tn, ea, i, tid_team, kFusedTeam, team_bar, red, phi_s, gbeg, masses, hb[0]);
                            }
                            if (tid_team == 0) {
                                const double s = __ldcg(a.score + i);
                                if (better(s, i, bs, bi)) { bs = s; bi = i; }
                            }
                        }
                    }
                } else {
                    // many rows: a warp per row
                    for (int64_t item = (int64_t)blockIdx.x * kFusedWarps + warp; item < n_items;
                         item += (int64_t)gridDim.x * kFusedWarps) {
                        const int64_t i = __ldcg(a.worklist + item);
                        if (tag_step(__ldcg(a.tags + i), a.epoch) != tn) {
                            catchup_row_call<XT>(i, lane, X, a.d, a.d_pad, recs, a.rec_len, a.w_cap, a.W, tn, a.sqn, a.U,
                                            a.ldu, a.tags, a.epoch, a.var, a.neg2ls2, uv);
                            eval_dispatch(tn, ea, i, lane, 32, 0, red, phi_s, gbeg, masses, hb[0]);
                        }
                        __syncwarp();
                        const double s = __ldcg(a.score + i);
                        if (better(s, i, bs, bi)) { bs = s; bi = i; }
                    }
                }
            }
            const Best b = group_argmax(bs, bi, tid_team, kFusedTeam, team_bar, ss + team * 16, si + team * 16);
            if (tid_team == 0) a.blk_best[gt] = b;
            if (warp_team == 0 && b.idx >= 0)
                prefetch_row<XT>(b.idx, lane, X, a.d_pad, a.U, a.ldu, a.W + tn, a.m, a.v, a.sqn, a.gain, a.tags);
        }
        FUSED_MARK();
        grid_barrier(a.barrier, target);
        FUSED_MARK();

        // ---- P6: the step's winner ---------------------------------------------------------------------------------
        {
            double bs = 0.0;
            long long bi = -1;
            for (int k = tid; k < n_teams; k += blockDim.x) {
                const double s = __ldcg(&a.blk_best[k].score);
                const long long i = __ldcg(&a.blk_best[k].idx);
                if (better(s, i, bs, bi)) { bs = s; bi = i; }
            }
            win = group_argmax(bs, bi, tid, kFusedThreads, 0, ss, si);
            if (blockIdx.x == 0 && tid < 4) {
                a.stats[4 * tn + tid] = __ldcg(a.counters + tid);
                if (tid < 3) a.counters[tid] = 0;                       // ready for the next greedy step
            }
        }
        FUSED_MARK();
    }
#undef FUSED_MARK
}

}  // namespace italk
