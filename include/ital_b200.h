/*
 * ital_b200 -- C ABI of the B200-native ITAL batch-selection path.
 *
 * One `ital_shard` owns a contiguous block of rows of the pool on one GPU and everything the greedy
 * batch construction needs about those rows (posterior moments, Cholesky projections, candidate mask).
 * A single-GPU learner is one shard holding all rows; a multi-GPU learner is one shard per process, and the
 * host moves the small fixed-size "point records" between them (one per greedy step).
 *
 * The reference is pure Python; there is no FFI in it.  Each entry point below names the reference code it
 * replaces (paths relative to /root/reference).  Plain pointers and sizes only; no torch types.  Every
 * function returns 0 on success and a negative ITAL_E* code on failure; ital_last_error() gives the text.
 * Calls on one shard must be serialised by the caller.  Host arrays are only read/written during the call.
 */
#ifndef ITAL_B200_H
#define ITAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ital_shard ital_shard;

enum { ITAL_OK = 0, ITAL_EINVAL = -1, ITAL_ECUDA = -2, ITAL_ESTATE = -3, ITAL_ENOMEM = -4 };
enum { ITAL_F32 = 0, ITAL_F64 = 1 };

/* Number of doubles in the header of a point record: [0] global row index, [1] score, [2] posterior mean,
 * [3] variance conditional on labelled + already selected points, [4] squared norm of the row,
 * [5] posterior variance given the labelled set, [6] gain (score - H(base)), [7] reserved.
 * The header is followed by `ital_width_cap()` projection entries and the row itself as `d` doubles. */
#define ITAL_RECORD_HEADER 8

const char* ital_last_error(void);
int ital_version(void);

/* GaussianProcess.__init__ (ital/gp.py:103-129) + ActiveRetrievalBase.fit (ital/retrieval_base.py:34-45),
 * without the n-by-n matrix: uploads rows [row_offset, row_offset + n_local) of the (data ++ queries) matrix,
 * computes their squared norms and resets the model.  `n_data` is the global number of pool rows; rows with a
 * global index >= n_data are query rows and never candidates.  `x_dtype` is the dtype of X AND of the copy
 * kept in HBM (ITAL_F32 halves the streamed bytes; use it only if the data is float32-representable). */
int ital_create(ital_shard** out, int device, const void* X, int x_dtype, int64_t n_local, int64_t d,
                int64_t row_offset, int64_t n_data, double length_scale, double var, double noise);
int ital_destroy(ital_shard* s);

/* Use this CUDA stream (a cudaStream_t) for all work of the shard; NULL = the legacy default stream.  A fetch also
 * uses one internal non-blocking stream (the scoring of the next greedy step runs beside the streaming pass of the
 * current one); that work is forked from and joined back into this stream with events inside the same call, so the
 * caller sees ordinary stream order. */
int ital_set_stream(ital_shard* s, void* cuda_stream);

/* GaussianProcess.reset (ital/gp.py:132-138) + ActiveRetrievalBase.reset (ital/retrieval_base.py:48-61):
 * forget all labels and all "seen" marks. */
int ital_reset(ital_shard* s);

/* Size in doubles of one point record in the shard's current state, and the projection capacity in it. */
int64_t ital_record_doubles(const ital_shard* s);
int64_t ital_width_cap(const ital_shard* s);
int64_t ital_width(const ital_shard* s);          /* labelled points currently in the model */

/* Fill host `records` (q records) for the given global rows.  Rows not owned by this shard give an all-zero
 * record, so that summing the buffers of all shards yields the complete records. */
int ital_export_points(ital_shard* s, int q, const int64_t* global_idx, double* records);

/* GaussianProcess.fit / update (ital/gp.py:141-200) + the predict_stored() that follows it
 * (ital/retrieval_base.py:58,120): add ONE labelled point (complete record, target y) to the model by a
 * rank-1 Cholesky extension and update posterior mean and variance of every local row in one pass over X.
 * Also marks the point as seen.  Points are appended in the order given (relevant first, irrelevant after,
 * as ActiveRetrievalBase.update does). */
int ital_add_labelled(ital_shard* s, const double* record, double y);
/* The same for q <= 4 points with ONE pass over X (block Cholesky extension; GaussianProcess.update with several
 * samples, ital/gp.py:164-200).  All q records must have been exported in the current state of the model, i.e.
 * before any of them is added. */
int ital_add_labelled_many(ital_shard* s, int q, const double* records, const double* y);

/* The same without the host in the loop, for a learner whose rows all live on this shard (one GPU): q <= 4 labelled
 * pool rows (global indices, targets y) enter the model with one small kernel that gathers their records, computes
 * the triangle of the block Cholesky extension among them and appends the model rows on the device
 * (k_prepare_labelled), followed by ONE pass over the pool.  Asynchronous: nothing waits for the GPU, nothing is read
 * back.  Marks the points as seen.  (ActiveRetrievalBase.update, ital/retrieval_base.py:105-126.) */
int ital_update_labelled(ital_shard* s, int q, const int64_t* global_idx, const double* y);

/* ActiveRetrievalBase.update's unnameable_ids / get_unseen (ital/retrieval_base.py:78-87,126):
 * mark rows as seen (never candidates again until reset).  Non-local indices are ignored. */
int ital_mark_seen(ital_shard* s, int64_t m, const int64_t* global_idx);

/* ITAL.fetch_unlabelled's top_candidates restriction (ital/ital.py:111-117): candidates are the unseen rows
 * whose global index is in `global_idx` (m entries; m < 0 lifts the restriction). */
int ital_restrict_candidates(ital_shard* s, int64_t m, const int64_t* global_idx);
/* The same restriction computed on the device for a learner on one shard: the `top` unseen pool rows with the largest
 * posterior mean stay candidates (np.argpartition(rel_mean[candidates], -top)[-top:], ital/ital.py:116; exact ties at
 * the cut go to the lower row).  A masked radix sort of the means; no n-sized copy to the host.  Lift with
 * ital_restrict_candidates(s, -1, NULL). */
int ital_restrict_top(ital_shard* s, int64_t top);

/* ---- greedy batch construction (ITAL.fetch_unlabelled, ital/ital.py:119-134) ---------------------------
 * ital_fetch_begin:   AppendedMutualInformation.__init__/set_ret([]) (ital/ital.py:491-558).
 * ital_fetch_propose: scores the local candidates of the current greedy step given the batch so far
 *                     (MutualInformation._call_iter_all, prob_rel; ital/ital.py:183-224, 345-383, with
 *                     predict_cov_batch, ital/gp.py:235-261, maintained incrementally) and writes the record
 *                     of the local best candidate (score desc, index asc; record[1] = -inf if none).
 *                     `floor_score`: a score some other shard already reached this step (or -inf).
 *                     `exhaustive` != 0 scores every candidate instead of pruning with the lazy-greedy
 *                     bound (same winner; used by tests and to report unpruned throughput).
 * ital_fetch_commit:  np.argmax + AppendedMutualInformation.append (ital/ital.py:130-132, 561-586) for the
 *                     globally best record: extends every local row's projection by the chosen point.
 * ital_fetch_end:     drops the batch-conditional state (the selected rows stay unseen until update()).
 * ital_fetch:         the whole loop for a single-shard learner; returns the number selected. */
/* Device-resident variants (what the multi-GPU host loop uses; nothing in them waits for the GPU):
 * ital_fetch_propose_dev writes the local best record to DEVICE memory (ital_record_doubles() doubles);
 * ital_fetch_commit_dev picks the best of `n_records` consecutive records in DEVICE memory (e.g. the output of
 * an NCCL all-gather of every shard's proposal) with the rule above, appends it to the batch and, if `extend`
 * is non-zero, runs the streaming pass (pass 0 for the last sample of the batch: nothing follows it);
 * ital_fetch_result synchronises and returns the (global row, score) of the samples committed so far. */
int ital_fetch_propose_dev(ital_shard* s, double floor_score, int exhaustive, double* record_dev);
int ital_fetch_commit_dev(ital_shard* s, const double* records_dev, int n_records, int extend);
int ital_fetch_result(ital_shard* s, int max_out, int64_t* out_idx, double* out_scores);
int ital_fetch_begin(ital_shard* s, double label_prob, double mistake_prob);
int ital_fetch_propose(ital_shard* s, double floor_score, int exhaustive, double* record);
int ital_fetch_commit(ital_shard* s, const double* record);
/* VarianceSampling.fetch_unlabelled (ital/baseline_methods.py:124-155) on the same batch state: inside a fetch
 * (ital_fetch_begin ... ital_fetch_end) score every local candidate by the change of "sum of variances minus sum of
 * covariances" of the batch when it is appended -- v_i - sum_a cov(r_a, i), from the incremental Cholesky rows the
 * streaming pass of ital_fetch_commit maintains (lazy rows must be off) -- or, with use_correlations == 0, by its
 * posterior variance alone, and write the record of the local best (score desc, row asc).  The reference's first pick
 * with use_correlations does not exclude unnameable rows (baseline_methods.py:133): first_pick_takes_unnameable != 0
 * mirrors that. */
int ital_variance_propose(ital_shard* s, int use_correlations, int first_pick_takes_unnameable, double* record);
/* ITAL(change_estimation_subset = c > 0) (ital/ital.py:102-108, AppendedMutualInformation.__call__ ital.py:514-527,
 * MutualInformation._call_iter_sub ital.py:227-275): the change of the model output is estimated on a random subset S
 * of the unseen samples kept at the signs of its means, while the relevance of the batch and the candidate is
 * integrated over.  Protocol (the host draws S with the reference's own np.random.choice call): ital_set_sub_mode(1),
 * lazy rows off; per evaluation ital_fetch_begin, ital_fetch_commit of the records of ext = [the n_batch samples picked
 * so far, then the members of S outside the batch] (every commit is one streaming pass, so each row holds its
 * projection on ext), then ital_fetch_propose_sub writes the record of the best local candidate outside ext -- or of
 * the single row only_row >= 0 (a member of S is scored with itself moved out of the subset) -- then ital_fetch_end.
 * Users who label every sample (label_prob = 1; mistake_prob of the fetch is honoured), label_estimation 'mean', at
 * most 7 batch samples and 11 columns. */
int ital_set_sub_mode(ital_shard* s, int on);
/* ITAL(clip_cov = th) (ital/ital.py:360-362, MutualInformation._grouped_prob_rel ital.py:386-429, group_cov
 * ital.py:590-616): from the sixth sample of a batch on, correlations of at most th are dropped and the orthant
 * probability factorises over the resulting groups.  For users who label every sample without mistakes the score of a
 * candidate is then the entropy of the batch's groups it is not connected to plus the entropy of the group it forms
 * with the ones it is; every candidate is scored at those steps (no lazy-greedy bound for the clipped model).
 * th outside (0, 1) turns it off. */
int ital_set_clip_cov(ital_shard* s, double clip_cov);
int ital_fetch_propose_sub(ital_shard* s, int n_batch, int64_t only_row, double* record);
int ital_fetch_end(ital_shard* s);
int ital_fetch(ital_shard* s, int k, double label_prob, double mistake_prob, int exhaustive,
               int64_t* out_idx, double* out_scores);

/* Multi-GPU fetch without NCCL in the loop.  Every shard owns an exchange buffer (one slot per shard and epoch
 * parity, one flag word per shard) that its peers map through CUDA IPC.  In a greedy step a shard stores its proposal
 * straight into the slot reserved for it in every peer's buffer (k_record: peer stores over NVLink / NVSwitch,
 * then a system-scope release of the epoch flag); the kernel that picks the winner (k_pick_winner) waits on the
 * flags of its own buffer.  This replaces the per-step all-gather of `ital_fetch_propose_dev` /
 * `ital_fetch_commit_dev` (np.argmax over the Pool's results, ital/ital.py:124-130, across GPUs).
 *   ital_peer_export:   (re)allocates the local buffer, writes its IPC handle (>= 64 bytes) to handle_out;
 *   ital_peer_connect:  maps the buffers of all shards; `handles` holds world handles of handle_bytes each, by rank;
 *   ital_fetch_peer:    the whole greedy loop (as ital_fetch); every shard must call it with the same arguments;
 *                       fails if a peer does not deliver within 5 s (bounded spin: the GPU is never left hanging);
 *   ital_peer_disconnect: unmaps the peers' buffers (call on every shard before any shard is destroyed);
 *   ital_peer_slot_doubles: capacity of a slot in doubles (0 if not connected); records longer than this
 *                       (more than 2048 projection entries) need the NCCL path. */
int ital_peer_export(ital_shard* s, int world, int rank, void* handle_out, int64_t handle_bytes);
int ital_peer_connect(ital_shard* s, const void* handles, int64_t handle_bytes);
int ital_peer_disconnect(ital_shard* s);
int64_t ital_peer_slot_doubles(const ital_shard* s);
int ital_fetch_peer(ital_shard* s, int k, double label_prob, double mistake_prob, int exhaustive, int64_t* out_idx,
                    double* out_scores);

/* Lazy rows (off by default).  Off: every greedy step streams the whole pool once to extend every row's
 * batch-conditional projection (k_extend; HBM-bound, what predict_cov_batch does for all rows, ital/gp.py:235-261).
 * On: the projection is extended only for the rows that are actually scored, on demand, from the stored records of
 * the selected points (k_catchup); same entries bit for bit, same batch, far less traffic when the lazy-greedy bound
 * prunes most rows. */
int ital_set_lazy_rows(ital_shard* s, int on);

/* ITAL(label_estimation=...) (ital/ital.py:62-67, 210-219): 0 'mean' (default: expectation over the relevance
 * configurations), 1 'optimistic' / 2 'pessimistic' (the largest / the "first or smaller" single term of the
 * enumeration, kept in the reference's order).  Other than 'mean' there is no lazy-greedy bound: every candidate is
 * scored at every step by the general-model kernel, batches of at most 5 samples. */
int ital_set_label_estimation(ital_shard* s, int mode);

/* The fused persistent fetch kernel (on by default; environment ITAL_B200_FUSED=0 turns it off).  With lazy rows on, a
 * user who labels every sample (label_prob >= 1) and pruning (exhaustive == 0), ital_fetch / ital_fetch_peer run the
 * first four greedy steps -- scores of the first step, quadrature nodes, stage A, worklist, exact scores, argmax and
 * commit of every step, with the peer exchange if there is one -- as ONE cooperative kernel with grid-wide barriers
 * (k_fetch_fused) instead of ~10 dependent launches per step; later steps of a longer batch continue with the
 * multi-kernel loop.  Same batch and bit-identical scores either way. */
int ital_set_fused(ital_shard* s, int on);

/* Diagnostics: with `on`, CTA 0 of k_fetch_fused stamps the GPU's nanosecond timer at every phase boundary; `out`
 * (may be NULL) receives up to max_out (<= 64) stamps of the last fused launch, in order: start, end of S0's scan,
 * after its barrier, commit of step 0, then per later step both sides of each of the five barriers (P1..P5) and the
 * commit.  Returns the number of stamps copied. */
int64_t ital_fused_trace(ital_shard* s, int on, uint64_t* out, int64_t max_out);

/* Streaming pass variant (on by default): stage the rows through shared memory with the bulk-copy engine (TMA,
 * cp.async.bulk + mbarrier; k_extend_bulk) where the tuned shape applies (2 KB rows, e.g. d = 512 float32), else
 * coalesced register loads (k_extend).  Same results bit for bit either way. */
int ital_set_bulk_stream(ital_shard* s, int on);

/* Per-step diagnostics of the last propose: [0] candidates considered, [1] candidates scored exactly,
 * [2] quadrature nodes, [3] H(base), [4] flagged (conditional variance < 100 * noise), [5] greedy steps the fused
 * kernel ran in the last ital_fetch / ital_fetch_peer, [6] quadrature nodes kept (device-generated rules drop the
 * nodes lighter than 1e-13), [7] reserved. */
int ital_fetch_stats(const ital_shard* s, double* out8);

/* Scores of the last propose for all local rows (NaN where not scored this step). */
int ital_last_scores(ital_shard* s, double* out_n_local);

/* rel_mean = gp.predict_stored()[:n] (ital/retrieval_base.py:58,120; ital/gp.py:221-222) and the posterior
 * variance of predict_stored(cov_mode='diag') before clamping (ital/gp.py:229), for the local rows. */
int ital_rel_mean(ital_shard* s, double* out_n_local);
int ital_rel_var(ital_shard* s, double* out_n_local);

/* ActiveRetrievalBase.top_results (ital/retrieval_base.py:64-75): np.argsort(rel_mean)[::-1][:k] over the local pool
 * rows, sorted on the device (stable radix sort; exactly tied means keep ascending row order, which the reference's
 * unstable argsort leaves undefined).  k < 0 = all local pool rows.  Writes global row indices (and, if out_val is
 * not NULL, their means) in descending order of the mean and returns how many were written; a multi-shard learner
 * merges the shards' lists on the host. */
int64_t ital_top_results(ital_shard* s, int64_t k, int64_t* out_idx, double* out_val);

/* GaussianProcess.predict (ital/gp.py:264-292): mean (and, if out_var != NULL, the 'diag' variance clamped at
 * 0) for m arbitrary rows of float64 features.  Uses the labelled rows held by this shard's model, so it is
 * valid on any shard (the labelled points are replicated). */
int ital_predict(ital_shard* s, const double* Xt, int64_t m, double* out_mean, double* out_var);
/* The same with the projections u = L_K^-1 k(X_L, x) of the m rows (out_proj[m][ital_width()], row-major), from which
 * the caller forms the 'full' covariance k(Xt, Xt) - U U^T of GaussianProcess.predict (ital/gp.py:285-287). */
int ital_predict_proj(ital_shard* s, const double* Xt, int64_t m, double* out_mean, double* out_proj);

/* Measurement hooks (bench.py): with profiling on, every launch of the streaming kernel (k_extend) is bracketed
 * by CUDA events on the shard's stream.  ital_profile_read synchronises, returns the accumulated kernel time in
 * milliseconds, the number of launches and the algorithmic bytes they moved, and clears the accumulators.
 * ital_launch_count: kernels launched by this shard since creation (all kernels, profiling on or off). */
int ital_profile_enable(ital_shard* s, int on);
int ital_profile_read(ital_shard* s, double* ms_total, int64_t* launches, double* algorithmic_bytes);
int64_t ital_launch_count(const ital_shard* s);
/* Bytes the library has copied host->device / device->host since the shard was created (every copy is counted where
 * it is issued; peer stores and NCCL traffic between GPUs are not host transfers). */
int ital_transfer_bytes(const ital_shard* s, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* Host-side pieces exposed for CPU-only tests (no GPU needed) ------------------------------------------- */
/* Shared quadrature nodes of one greedy step (see oracle/orthant.py for the rule): base mean m[t], lower
 * Cholesky factor L[t*t] (row-major).  With eta == NULL returns the capacity ((2q)^t, or 65536 quasi-Monte-Carlo
 * nodes from t = 4 on); otherwise fills eta[t*N]
 * (dimension-major, stride N), w[N], orth[N], sorted by orthant, and masses[2^t], and returns N, the number of
 * nodes kept (nodes lighter than 1e-13 are dropped). */
int64_t ital_snq_nodes(int t, const double* m, const double* L, double* eta, double* w, int32_t* orth,
                       double* masses);
int ital_snq_order(int t);
/* Table behind the closed-form score of the first greedy step (one sample: ital/ital.py:364-369 + 183-224),
 *   H(u) = -Phi(u) log(Phi(u) + 1e-12) - Phi(-u) log(Phi(-u) + 1e-12),  u = |mean| / stdev in [0, 8.5]:
 * 136 intervals of width 1/16, 8 coefficients each (ascending powers of x = 32 (u - (k + 1/2) / 16) in [-1, 1]).
 * Copies up to max_out doubles, returns the table length (1088). */
int64_t ital_h_table(double* out, int64_t max_out);
/* Table behind the standard normal CDF of the scoring kernels (replaces scipy.stats.norm.cdf / the CDF inside mvndst,
 * ital/ital.py:364-383): (Phi, phi) at the 2177 grid points x_k = -8.5 + k/128; the kernels add a 5th-order Taylor
 * step from the nearest grid point (phi_tab in csrc/ital_kernels.cuh).  Copies up to max_out doubles, returns the
 * length (4354). */
int64_t ital_phi_table(double* out, int64_t max_out);
/* Node sets and conditional moments of ital_fetch_propose_sub (change_estimation_subset; csrc/snq_host.h
 * generate_sub) for ext = [n_batch batch samples, n_cols - n_batch subset members] with means m[n_cols] and row-major
 * Cholesky factor L[n_cols^2].  sizes[4] = {n_nodes, n_groups = 2 * 2^n_batch, s* (sign bits of the subset's means),
 * doubles in `tables`}; call with eta == NULL to get the sizes only.  eta[n_cols * n_nodes] dimension-major,
 * w[n_nodes], group_begin[n_groups + 1]; tables = mass[2 G] | mu[G D] | Sig[D D] | mU[G u] | CU[u u] | BS[u D]. */
int ital_snq_sub(int n_batch, int n_cols, const double* m, const double* L, double noise, int64_t* sizes, double* eta,
                 double* w, int32_t* group_begin, double* tables);
/* Conditional node sets of the general feedback model (label_prob < 1; csrc/snq_host.h generate_general).
 * sizes[4] = {n_nodes, n_groups, n_sets, lut entries}; call with eta == NULL to get the sizes only.  eta[t*n_nodes]
 * dimension-major, w[n_nodes], group_begin[n_groups+1], group_mass[n_groups], set_group0[n_sets+1],
 * lut[3 * 4^(t+1)] = {group, set, flags} per (relevance configuration, labelled subset). */
int ital_snq_general(int t, const double* m, const double* L, double noise, int64_t* sizes, double* eta, double* w,
                     int32_t* group_begin, double* group_mass, int32_t* set_group0, int32_t* lut);

#ifdef __cplusplus
}
#endif
#endif
