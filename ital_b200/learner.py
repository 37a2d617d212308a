"""`ITAL` learner with the reference's public interface over the CUDA library (include/ital_b200.h).

Mirrors /root/reference/ital/ital.py:12-134 (ITAL) and /root/reference/ital/retrieval_base.py:7-194
(ActiveRetrievalBase): same constructor keywords, same methods, same attributes read by the reference's
`run_experiment.py` / `utils.LEARNERS`, same exceptions.  Everything that scales with the pool runs on the
GPU; there is no CPU fallback (a missing library or device raises).
"""
import ctypes

import numpy as np

from . import _capi
from .dist import LocalComm, TorchComm, merge_top, partition_rows, pick_winner


class _Shard(object):
    """Thin owner of one `ital_shard*`."""

    def __init__(self, X, dtype_code, row_offset, n_data, length_scale, var, noise, device):
        self.lib = _capi.load()
        self.handle = ctypes.c_void_p()
        self.n_local, self.d = X.shape
        self._keepalive = X
        _capi.check(self.lib.ital_create(ctypes.byref(self.handle), int(device), X.ctypes.data_as(ctypes.c_void_p),
                                         dtype_code, self.n_local, self.d, int(row_offset), int(n_data),
                                         float(length_scale), float(var), float(noise)))
        self._keepalive = None

    def close(self):
        if self.handle:
            self.lib.ital_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    __del__ = close

    def record_doubles(self):
        return int(self.lib.ital_record_doubles(self.handle))

    def reset(self):
        _capi.check(self.lib.ital_reset(self.handle))

    def set_stream(self, stream_ptr):
        _capi.check(self.lib.ital_set_stream(self.handle, ctypes.c_void_p(stream_ptr)))

    def export_points(self, idx):
        idx = _capi.as_i64(idx)
        rec = np.zeros((len(idx), self.record_doubles()))
        _capi.check(self.lib.ital_export_points(self.handle, len(idx), _capi.i64ptr(idx), _capi.dptr(rec)))
        return rec

    def add_labelled(self, record, y):
        record = _capi.as_f64(record)
        _capi.check(self.lib.ital_add_labelled(self.handle, _capi.dptr(record), float(y)))

    def add_labelled_many(self, records, y):
        records, y = _capi.as_f64(records), _capi.as_f64(y)
        _capi.check(self.lib.ital_add_labelled_many(self.handle, len(y), _capi.dptr(records), _capi.dptr(y)))

    def update_labelled(self, idx, y):
        idx, y = _capi.as_i64(idx), _capi.as_f64(y)
        _capi.check(self.lib.ital_update_labelled(self.handle, len(idx), _capi.i64ptr(idx), _capi.dptr(y)))

    def mark_seen(self, idx):
        idx = _capi.as_i64(idx)
        if len(idx):
            _capi.check(self.lib.ital_mark_seen(self.handle, len(idx), _capi.i64ptr(idx)))

    def restrict_candidates(self, idx):
        if idx is None:
            _capi.check(self.lib.ital_restrict_candidates(self.handle, -1, None))
        else:
            idx = _capi.as_i64(idx)
            _capi.check(self.lib.ital_restrict_candidates(self.handle, len(idx), _capi.i64ptr(idx)))

    def restrict_top(self, top):
        _capi.check(self.lib.ital_restrict_top(self.handle, int(top)))

    def fetch_begin(self, label_prob, mistake_prob):
        _capi.check(self.lib.ital_fetch_begin(self.handle, float(label_prob), float(mistake_prob)))

    def fetch_propose(self, floor_score, exhaustive):
        rec = np.zeros(self.record_doubles())
        _capi.check(self.lib.ital_fetch_propose(self.handle, float(floor_score), int(bool(exhaustive)),
                                                _capi.dptr(rec)))
        return rec

    def variance_propose(self, use_correlations, first_pick_takes_unnameable):
        rec = np.zeros(self.record_doubles())
        _capi.check(self.lib.ital_variance_propose(self.handle, int(bool(use_correlations)),
                                                   int(bool(first_pick_takes_unnameable)), _capi.dptr(rec)))
        return rec

    def set_sub_mode(self, on):
        _capi.check(self.lib.ital_set_sub_mode(self.handle, int(bool(on))))

    def fetch_propose_sub(self, n_batch, only_row):
        rec = np.zeros(self.record_doubles())
        _capi.check(self.lib.ital_fetch_propose_sub(self.handle, int(n_batch), int(only_row), _capi.dptr(rec)))
        return rec

    def fetch_commit(self, record):
        record = _capi.as_f64(record)
        _capi.check(self.lib.ital_fetch_commit(self.handle, _capi.dptr(record)))

    def fetch_end(self):
        _capi.check(self.lib.ital_fetch_end(self.handle))

    def fetch_propose_dev(self, floor_score, exhaustive, dev_ptr):
        _capi.check(self.lib.ital_fetch_propose_dev(self.handle, float(floor_score), int(bool(exhaustive)),
                                                    ctypes.c_void_p(dev_ptr)))

    def fetch_commit_dev(self, dev_ptr, n_records, extend):
        _capi.check(self.lib.ital_fetch_commit_dev(self.handle, ctypes.c_void_p(dev_ptr), int(n_records),
                                                   int(bool(extend))))

    def fetch_result(self, k):
        idx = np.zeros(max(k, 1), dtype=np.int64)
        scores = np.zeros(max(k, 1))
        got = _capi.check(self.lib.ital_fetch_result(self.handle, int(k), _capi.i64ptr(idx), _capi.dptr(scores)))
        return idx[:got], scores[:got]

    def fetch(self, k, label_prob, mistake_prob, exhaustive):
        idx = np.zeros(max(k, 1), dtype=np.int64)
        scores = np.zeros(max(k, 1))
        got = _capi.check(self.lib.ital_fetch(self.handle, int(k), float(label_prob), float(mistake_prob),
                                              int(bool(exhaustive)), _capi.i64ptr(idx), _capi.dptr(scores)))
        return idx[:got], scores[:got]

    def peer_export(self, world, rank, handle_bytes):
        handle = np.zeros(handle_bytes, dtype=np.uint8)
        _capi.check(self.lib.ital_peer_export(self.handle, int(world), int(rank),
                                              handle.ctypes.data_as(ctypes.c_void_p), int(handle_bytes)))
        return handle

    def peer_connect(self, handles, handle_bytes):
        """True if the exchange buffers of all shards could be mapped (same node, peer access)."""
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        return self.lib.ital_peer_connect(self.handle, handles.ctypes.data_as(ctypes.c_void_p), int(handle_bytes)) == 0

    def peer_disconnect(self):
        if self.handle:
            self.lib.ital_peer_disconnect(self.handle)

    def peer_slot_doubles(self):
        return int(self.lib.ital_peer_slot_doubles(self.handle))

    def fetch_peer(self, k, label_prob, mistake_prob, exhaustive):
        idx = np.zeros(max(k, 1), dtype=np.int64)
        scores = np.zeros(max(k, 1))
        got = _capi.check(self.lib.ital_fetch_peer(self.handle, int(k), float(label_prob), float(mistake_prob),
                                                   int(bool(exhaustive)), _capi.i64ptr(idx), _capi.dptr(scores)))
        return idx[:got], scores[:got]

    def stats(self):
        out = np.zeros(8)
        _capi.check(self.lib.ital_fetch_stats(self.handle, _capi.dptr(out)))
        return out

    def _vec(self, fn):
        out = np.zeros(self.n_local)
        _capi.check(fn(self.handle, _capi.dptr(out)))
        return out

    def last_scores(self):
        return self._vec(self.lib.ital_last_scores)

    def rel_mean(self):
        return self._vec(self.lib.ital_rel_mean)

    def rel_var(self):
        return self._vec(self.lib.ital_rel_var)

    def top_results(self, k):
        cap = self.n_local if k is None or k < 0 else min(int(k), self.n_local)
        idx = np.zeros(max(cap, 1), dtype=np.int64)
        val = np.zeros(max(cap, 1))
        got = _capi.check(self.lib.ital_top_results(self.handle, -1 if k is None else int(k), _capi.i64ptr(idx),
                                                    _capi.dptr(val)))
        return idx[:got], val[:got]

    def predict_proj(self, X):
        X = _capi.as_f64(X)
        W = int(self.lib.ital_width(self.handle))
        mean, proj = np.zeros(len(X)), np.zeros((len(X), W))
        _capi.check(self.lib.ital_predict_proj(self.handle, _capi.dptr(X), len(X), _capi.dptr(mean), _capi.dptr(proj)))
        return mean, proj

    def predict(self, X, want_var):
        X = _capi.as_f64(X)
        mean = np.zeros(len(X))
        var = np.zeros(len(X)) if want_var else None
        _capi.check(self.lib.ital_predict(self.handle, _capi.dptr(X), len(X), _capi.dptr(mean),
                                          _capi.dptr(var) if want_var else None))
        return mean, var


class _GPView(object):
    """What the reference's callers read from `learner.gp` (run_experiment.py:153,166; viz_utils.py:153,177)."""

    def __init__(self, learner):
        self._l = learner

    @property
    def ind(self):
        return list(self._l._labelled_idx)

    @property
    def y(self):
        return np.array(self._l._labelled_y, dtype=np.float64)

    @property
    def noise(self):
        return self._l.noise

    @property
    def var(self):
        return self._l.var

    @property
    def length_scale(self):
        return self._l.length_scale

    def predict(self, X, cov_mode=None):
        """GaussianProcess.predict (ital/gp.py:264-292); cov_mode None, 'diag' or 'full'."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        if cov_mode == 'full':                  # k(X, X) - k^T K^-1 k (gp.py:285-287) from the rows' projections
            mean, proj = self._l._shard.predict_proj(X)
            sq = np.sum(X ** 2, axis=-1)
            kxx = self._l.var * np.exp((sq[:, None] + sq[None, :] - 2.0 * (X @ X.T)) / (-2.0 * self._l.length_scale ** 2))
            return mean, kxx - proj @ proj.T
        mean, var = self._l._shard.predict(X, cov_mode == 'diag')
        return (mean, var) if cov_mode == 'diag' else mean

    def predict_stored(self, ind=None, cov_mode=None):
        """GaussianProcess.predict_stored (ital/gp.py:203-232) for the pool rows; cov_mode None or 'diag'."""
        if cov_mode == 'full':
            if ind is None:
                raise NotImplementedError("cov_mode='full' over all rows is the n-by-n matrix this path never forms")
            return self._l._posterior_block(ind)
        mean = self._l._all_rows(self._l._shard.rel_mean())
        sel = slice(None) if ind is None else np.asarray(ind, dtype=np.int64)
        if cov_mode == 'diag':
            return mean[sel], np.maximum(0, self._l._all_rows(self._l._shard.rel_var()))[sel]
        return mean[sel]

    def __getattr__(self, name):
        if name in ('K_all', 'K_inv', 'K', 'w'):
            raise AttributeError("gp.%s does not exist here: the n-by-n kernel matrix of ital/gp.py:128 is never "
                                 "formed and the model is kept as a Cholesky factor" % name)
        raise AttributeError(name)


class ITAL(object):
    """Information-theoretic Active Learning; drop-in for `ital.ITAL` (ital/ital.py:12-134).

    Extra keywords (all optional): `device` CUDA ordinal; `storage` 'auto' | 'float32' | 'float64' for the copy
    of the data kept in HBM ('auto' picks float32 only if that is lossless); `process_group` a
    torch.distributed group (or True for the default group) to shard the rows over one GPU per process;
    `exhaustive` scores every candidate each step instead of pruning with the lazy-greedy bound;
    `local_rows=(first_row, n_total)` declares that `data` holds only this process's contiguous block of a
    pool of n_total rows (for pools too large to replicate on every host process; no `queries` then);
    `lazy_rows` extends the batch-conditional projections only for the rows that get scored instead of streaming
    the whole pool once per greedy step (same batch, same scores bit for bit; see include/ital_b200.h); the default
    (None) does so whenever the lazy-greedy bound prunes, i.e. for users who label every sample unless `exhaustive`;
    `bulk_stream` (default on; environment ITAL_B200_BULK=0 turns it off) stages the streaming pass through shared
    memory with the bulk-copy engine where the tuned shape applies (2 KB rows), else coalesced register loads.
    """

    def __init__(self, data=None, queries=[], length_scale=0.1, var=1.0, noise=1e-6,
                 label_prob=1.0, mistake_prob=0.0, top_candidates=None, change_estimation_subset=0,
                 clip_cov=0, label_estimation='mean', monte_carlo_num_rel=None, monte_carlo_num_fb=None,
                 parallelized=True, device=None, storage='auto', process_group=None, exhaustive=False,
                 local_rows=None, lazy_rows=None, bulk_stream=None):
        self.length_scale, self.var, self.noise = length_scale, var, noise
        self.label_prob, self.mistake_prob = label_prob, mistake_prob
        self.top_candidates = top_candidates
        self.change_estimation_subset = change_estimation_subset
        self.clip_cov = clip_cov
        self.label_estimation = label_estimation
        self.monte_carlo_num_rel, self.monte_carlo_num_fb = monte_carlo_num_rel, monte_carlo_num_fb
        self.parallelized = parallelized            # accepted for compatibility; the GPU is the parallelism
        self.exhaustive = exhaustive
        self._lazy_rows = None if lazy_rows is None else bool(lazy_rows)
        import os
        self._bulk_stream = os.environ.get('ITAL_B200_BULK', '1') == '1' if bulk_stream is None else bool(bulk_stream)
        self._fused = os.environ.get('ITAL_B200_FUSED', '1') == '1'
        self._storage = storage
        self._device = device
        self._local_rows = local_rows
        self._comm = LocalComm() if process_group is None else \
            TorchComm(None if process_group is True else process_group, device)
        self._shard = None
        self._peer = False
        self.last_fetch_stats = []
        self.last_fused_steps = 0
        self.fit(data, queries)

    @property
    def lazy_rows(self):
        return self._lazy_rows

    @lazy_rows.setter
    def lazy_rows(self, on):
        self._lazy_rows = None if on is None else bool(on)

    def _apply_lazy_rows(self):
        """None (default) = on demand whenever the lazy-greedy bound prunes (users who label every sample, not
        `exhaustive`): a greedy step then scores a few hundred rows and a pass over the whole pool per step would feed
        nothing; otherwise the streaming pass keeps every row's projection current (every row is scored)."""
        on = self._lazy_rows
        if on is None:
            on = self.label_prob >= 1 and not self._exhaustive()
        _capi.check(self._shard.lib.ital_set_lazy_rows(self._shard.handle, int(bool(on))))
        _capi.check(self._shard.lib.ital_set_label_estimation(self._shard.handle,
                                                              self.ESTIMATIONS.get(self.label_estimation, 0)))
        return bool(on)

    def _apply_clip_cov(self):
        clip = float(self.clip_cov) if self.clip_cov and 0 < self.clip_cov < 1 else 0.0     # ital.py:360
        _capi.check(self._shard.lib.ital_set_clip_cov(self._shard.handle, clip))

    def _exhaustive(self):
        """Every candidate is scored at every step: asked for, or no lazy-greedy bound (only the expectation over the
        relevance configurations, label_estimation='mean', is a submodular entropy)."""
        return bool(self.exhaustive) or self.label_estimation != 'mean'


    @property
    def fused(self):
        return self._fused

    @fused.setter
    def fused(self, on):
        self._fused = bool(on)
        if self._shard is not None:
            _capi.check(self._shard.lib.ital_set_fused(self._shard.handle, int(self._fused)))

    @property
    def bulk_stream(self):
        return self._bulk_stream

    @bulk_stream.setter
    def bulk_stream(self, on):
        self._bulk_stream = bool(on)
        if self._shard is not None:
            _capi.check(self._shard.lib.ital_set_bulk_stream(self._shard.handle, int(self._bulk_stream)))

    # ---- ActiveRetrievalBase ---------------------------------------------------------------------------
    def fit(self, data, queries=[]):                                            # retrieval_base.py:34-45
        self.data = data
        self.queries = queries
        self.close()
        if self.data is None:
            self.gp = None
            return
        X = np.asarray(self.data)
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        if len(self.queries) > 0:
            X = np.concatenate((X, np.asarray(self.queries, dtype=np.float64).reshape(len(self.queries), -1)))
        self._n = len(self.data) if self._local_rows is None else int(self._local_rows[1])
        if self._local_rows is not None and len(self.queries) > 0:
            raise ValueError('local_rows and queries cannot be combined')
        storage = self._storage
        if storage == 'auto':
            storage = 'float32' if X.dtype == np.float32 or np.array_equal(X.astype(np.float32), X) else 'float64'
        if storage not in ('float32', 'float64'):
            raise ValueError("storage must be 'auto', 'float32' or 'float64'")
        self.storage = storage
        if self._local_rows is None:
            self._offsets = partition_rows(len(X), self._comm.world_size)
            lo, hi = int(self._offsets[self._comm.rank]), int(self._offsets[self._comm.rank + 1])
            if hi <= lo:
                raise ValueError('more processes than rows')
            Xl = X[lo:hi]
        else:
            lo = int(self._local_rows[0])
            firsts = self._comm.gather_records(np.array([float(lo)]))[:, 0].astype(np.int64)
            self._offsets = np.concatenate((firsts, [self._n])).astype(np.int64)
            if np.any(np.diff(self._offsets) <= 0) or self._offsets[self._comm.rank + 1] - lo != len(X):
                raise ValueError('local_rows blocks must be contiguous, ordered by rank and cover the pool')
            Xl = X
        Xl = np.ascontiguousarray(Xl, dtype=np.float32 if storage == 'float32' else np.float64)
        device = self._device
        if device is None:
            device = getattr(self._comm, 'device', None)
            device = device.index if device is not None and getattr(device, 'type', 'cpu') == 'cuda' else 0
        self._shard = _Shard(Xl, _capi.ITAL_F32 if storage == 'float32' else _capi.ITAL_F64, lo, self._n,
                             self.length_scale, self.var, self.noise, device)
        self.gp = _GPView(self)
        self._setup_peer_exchange()
        self.bulk_stream = self._bulk_stream
        self.fused = self._fused
        self.reset()

    def _setup_peer_exchange(self):
        """Multi-GPU on one node: map every shard's exchange buffer into every process (CUDA IPC), so that a greedy
        step exchanges its proposals by peer stores over NVLink instead of an NCCL all-gather (ital_fetch_peer).
        Falls back to the NCCL loop, on all ranks alike, if any mapping fails (ITAL_B200_PEER=0 forces that)."""
        import os
        self._peer = False
        comm = self._comm
        if comm.world_size == 1 or not getattr(comm, 'on_device', False) \
                or os.environ.get('ITAL_B200_PEER', '1') != '1':
            return
        handle_bytes = 128
        try:                        # (more than 32 shards, no IPC support, ...: every rank falls back together)
            mine, ok = self._shard.peer_export(comm.world_size, comm.rank, handle_bytes), True
        except _capi.ItalError:
            mine, ok = np.zeros(handle_bytes, dtype=np.uint8), False
        handles = comm.gather_bytes(mine)
        ok = comm.all_agree(ok) and self._shard.peer_connect(handles, handle_bytes)
        self._peer = comm.all_agree(ok)
        if not self._peer:
            self._shard.peer_disconnect()

    def close(self):
        """Release the GPU state.  With several GPUs every process must call it (the peers' exchange buffers are
        unmapped between two barriers, before any of them is freed)."""
        if getattr(self, '_shard', None) is None:
            return
        if getattr(self, '_peer', False):
            self._comm.barrier()
            self._shard.peer_disconnect()
            self._comm.barrier()
            self._peer = False
        self._shard.close()
        self._shard = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        """Last resort (prefer close() / a with-block): non-collective release, no barrier -- a learner that is
        garbage-collected on one rank cannot take part in a collective fetch any more anyway.  The peers' mappings of
        this shard's exchange buffer are theirs to close."""
        shard = getattr(self, '_shard', None)
        if shard is not None:
            try:
                if getattr(self, '_peer', False):
                    shard.peer_disconnect()
                shard.close()
            except Exception:
                pass
            self._shard = None

    def reset(self):                                                            # retrieval_base.py:48-61
        self.rounds = 0
        self.relevant_ids, self.irrelevant_ids, self.unnameable_ids = set(), set(), set()
        self._labelled_idx, self._labelled_y = [], []
        self._rel_mean = None
        self._shard.reset()
        if len(self.queries) > 0:
            self._add_labelled(list(range(self._n, self._n + len(self.queries))), [1.0] * len(self.queries))

    @property
    def rel_mean(self):
        """gp.predict_stored()[:n] (retrieval_base.py:58,120); None before the first label."""
        if len(self._labelled_idx) == 0:
            return None
        if self._rel_mean is None:
            self._rel_mean = self._all_rows(self._shard.rel_mean())[:self._n]
        return self._rel_mean

    def _all_rows(self, local):
        return self._comm.gather_rows(local, self._offsets)

    def top_results(self, k=None):                                              # retrieval_base.py:64-75
        """np.argsort(rel_mean)[::-1][:k], sorted on the GPU (ital_top_results): only the k indices come back to the
        host.  Exactly tied means come in ascending row order (the reference's unstable sort leaves that open)."""
        if len(self._labelled_idx) == 0:
            raise RuntimeError('top_results() needs at least one query or labelled sample')
        want = None if k is None else (max(0, self._n + int(k)) if k < 0 else min(int(k), self._n))   # ind[:k]
        if want == 0:
            return np.zeros(0, dtype=np.int64)
        idx, val = self._shard.top_results(want)
        if self._comm.world_size == 1:
            return idx
        # every shard contributes its own top list; merge by (mean desc, row asc)
        widest = int(np.max(np.diff(self._offsets)))
        return merge_top(self._comm, idx, val, widest if want is None else min(want, widest), want)

    def _seen_mask(self):
        seen = np.zeros(self._n, dtype=bool)
        for ids in (self.relevant_ids, self.irrelevant_ids, self.unnameable_ids):
            if ids:
                seen[np.fromiter(ids, dtype=np.int64, count=len(ids))] = True
        return seen

    def get_unseen(self):                                                       # retrieval_base.py:78-87
        return np.nonzero(~self._seen_mask())[0].tolist()

    def partition_feedback(self, feedback):                                     # retrieval_base.py:167-194
        rel, irr, unnameable = [], [], []
        for i, fb in feedback.items():
            if fb > 0:
                if i in self.irrelevant_ids:
                    raise RuntimeError('Cannot change feedback once given.')
                elif i not in self.relevant_ids:
                    rel.append(i)
            elif fb < 0:
                if i in self.relevant_ids:
                    raise RuntimeError('Cannot change feedback once given.')
                elif i not in self.irrelevant_ids:
                    irr.append(i)
            else:
                unnameable.append(i)
        return rel, irr, unnameable

    def _add_labelled(self, idx, y):
        """gp.fit / gp.update (ital/gp.py:141-200): up to four labelled points per pass over the pool."""
        self._check_rows(idx, with_queries=True)
        idx, y = [int(i) for i in idx], [float(v) for v in y]
        for lo in range(0, len(idx), 4):
            chunk = idx[lo:lo + 4]
            if self._comm.world_size == 1:
                self._shard.update_labelled(chunk, y[lo:lo + 4])                # all on the device, nothing waits
            else:
                recs = self._comm.sum_records(self._shard.export_points(chunk))     # all in the current state
                self._shard.add_labelled_many(recs, y[lo:lo + 4])
            self._labelled_idx.extend(chunk)
            self._labelled_y.extend(y[lo:lo + 4])
        self._rel_mean = None

    def _check_rows(self, idx, with_queries=False):
        """The reference indexes numpy arrays with these (IndexError when out of range); a row no shard owns would
        otherwise enter the model as an all-zero record."""
        hi = self._n + (len(self.queries) if with_queries else 0)
        for i in idx:
            if not (0 <= int(i) < hi) or int(i) != i:
                raise IndexError('sample index %r is out of range for a pool of %d rows' % (i, self._n))

    def update(self, feedback):                                                 # retrieval_base.py:105-126
        self._check_rows(feedback.keys())
        rel, irr, unnameable = self.partition_feedback(feedback)
        if len(rel) + len(irr) > 0:
            self._add_labelled(rel + irr, [1.0] * len(rel) + [-1.0] * len(irr))
            self.relevant_ids.update(rel)
            self.irrelevant_ids.update(irr)
            self.rounds += 1
        if len(unnameable):
            self._shard.mark_seen([int(i) for i in unnameable])
        self.unnameable_ids.update(unnameable)

    def _posterior_block(self, rows):
        """Posterior means and full covariance of a small set of rows from their point records (one k_record launch
        per row): c_ij = k(x_i, x_j) - u_i . u_j with u the projections on the labelled set's Cholesky factor."""
        self._check_rows(rows, with_queries=True)
        rows = [int(i) for i in rows]
        recs = []
        for lo in range(0, len(rows), 64):
            recs.append(self._comm.sum_records(self._shard.export_points(rows[lo:lo + 64])))
        recs = np.concatenate(recs) if recs else np.zeros((0, self._shard.record_doubles()))
        W = int(self._shard.lib.ital_width(self._shard.handle))
        cap = int(self._shard.lib.ital_width_cap(self._shard.handle))
        h = _capi.RECORD_HEADER
        u, x = recs[:, h:h + W], recs[:, h + cap:h + cap + self._shard.d]
        sq = recs[:, 4]
        kxx = self.var * np.exp((sq[:, None] + sq[None, :] - 2.0 * (x @ x.T)) / (-2.0 * self.length_scale ** 2))
        return recs[:, 2].copy(), kxx - u @ u.T

    def updated_prediction(self, feedback, test_ind, cov_mode='full'):          # retrieval_base.py:129-164
        """Prediction for `test_ind` after hypothetically adding `feedback` (gp.py:295-344), without changing the
        model.  Meant for the handful of rows the reference calls it with (a batch and its candidates): the
        posterior block of the rows involved is assembled from their point records and conditioned on the
        annotated ones (mean' = m_T + C_TO (C_OO + noise I)^-1 (y - m_O), cov' = C_TT - C_TO (C_OO + noise I)^-1 C_OT),
        which equals the reference's extended inverse (extend_inv, gp.py:40-87)."""
        if cov_mode not in (None, 'diag', 'full'):
            raise ValueError('cov_mode must be None, "diag" or "full"')
        self._check_rows(feedback.keys())
        rel, irr, _ = self.partition_feedback(feedback)
        test_ind = [int(i) for i in test_ind]
        obs = sorted(rel) + sorted(irr)
        y = np.concatenate((np.ones(len(rel)), -np.ones(len(irr))))
        mean, cov = self._posterior_block(test_ind + obs)
        T, O = slice(0, len(test_ind)), slice(len(test_ind), len(test_ind) + len(obs))
        mean_t, cov_t = mean[T], cov[T, T]
        if len(obs):
            G = np.linalg.solve(cov[O, O] + self.noise * np.eye(len(obs)), cov[O, T])
            mean_t = mean_t + G.T @ (y - mean[O])
            cov_t = cov_t - cov[T, O] @ G
        if cov_mode == 'full':
            return mean_t, cov_t
        if cov_mode == 'diag':
            return mean_t, np.maximum(0, np.diag(cov_t))
        return mean_t

    # ---- ITAL ------------------------------------------------------------------------------------------
    ESTIMATIONS = {'mean': 0, 'optimistic': 1, 'pessimistic': 2}
    MAX_BATCH = 11                  # greedy steps per fetch (10 base variables)
    MAX_BATCH_GENERAL = 5           # with label_prob < 1 (conditional node sets up to 4 base variables)
    MAX_BATCH_SUBSET = 7            # with change_estimation_subset > 0 ...
    MAX_COLS_SUBSET = 11            # ... and batch - 1 + subset columns at most

    def _check_supported(self, k=0):
        if self.label_estimation not in self.ESTIMATIONS:
            raise ValueError("label_estimation must be 'mean', 'optimistic' or 'pessimistic'")
        general = self.label_prob < 1 or self.label_estimation != 'mean'
        limit = self.MAX_BATCH_GENERAL if general else self.MAX_BATCH
        if k > limit:               # before any work (and before any collective) -- not in the middle of the greedy loop
            raise NotImplementedError('batches of more than %d samples are not supported%s' % (
                limit, " with label_prob < 1 or label_estimation other than 'mean'" if general else ''))
        # clip_cov only matters from the sixth sample on (ital.py:360), i.e. only for label_prob = 1 and 'mean' (the other
        # models stop at 5 samples above); a user who mislabels adds the same per-step constant as without clipping
        ce = self.change_estimation_subset
        if ce is None:              # every unseen sample in the subset: orthant probabilities in n dimensions
            raise NotImplementedError('change_estimation_subset=None (all unseen samples) is not built')
        if ce > 0:
            if general:
                raise NotImplementedError("change_estimation_subset is built for label_prob=1, label_estimation='mean'")
            if k > self.MAX_BATCH_SUBSET or k - 1 + ce > self.MAX_COLS_SUBSET:
                raise NotImplementedError('change_estimation_subset: at most %d samples per batch and batch + subset '
                                          '<= %d' % (self.MAX_BATCH_SUBSET, self.MAX_COLS_SUBSET + 1))
        # monte_carlo_num_rel / monte_carlo_num_fb (ital.py:293-297, 319-342): the reference replaces the enumeration
        # of relevance / feedback configurations by random samples only when there are more configurations than
        # samples (2^(D-1) >= D * num, 3^D >= 2 * D * num); its samples come from the global numpy RNG in candidate
        # order, so its noisy estimates are not reproducible outside that exact call sequence.  Here the sums those
        # samples estimate are always evaluated exactly (the enumeration is cheap on the device): the zero-variance
        # limit of the reference's estimator.

    def fetch_unlabelled(self, k, show_progress=False):                         # ital.py:84-134
        """Greedy batch of k unlabelled samples maximising mutual information; list of row indices."""
        if len(self._labelled_idx) == 0:
            raise RuntimeError('fetch_unlabelled() needs at least one query or labelled sample '
                               '(the reference fails here with gp.K_inv = None)')
        n_unseen = self._n - len(self.relevant_ids | self.irrelevant_ids | self.unnameable_ids)
        k = min(int(k), n_unseen)                                               # ital.py:99-100
        self._check_supported(k)
        restricted = False
        if self.top_candidates is not None:                                     # ital.py:111-117
            top = self.top_candidates
            if isinstance(top, float):
                top = min(n_unseen, int(top * (len(self.queries) + len(self.relevant_ids) + len(self.irrelevant_ids))))
            if 0 < top < n_unseen:
                if self._comm.world_size == 1:
                    self._shard.restrict_top(top)           # masked sort of the means on the device
                else:                                       # (several shards: the cut is global)
                    cand = np.nonzero(~self._seen_mask())[0]
                    top_ind = np.argpartition(self.rel_mean[cand], -top)[-top:]
                    self._shard.restrict_candidates(cand[top_ind])
                restricted = True
        self.last_fetch_stats = []
        self._apply_lazy_rows()
        self._apply_clip_cov()
        try:
            if self.change_estimation_subset:
                return self._fetch_change_subset(k)
            if self._comm.world_size == 1 and not show_progress:
                idx, scores = self._shard.fetch(k, self.label_prob, self.mistake_prob, self._exhaustive())
                self.last_fetch_scores = scores
                self.last_fused_steps = int(self._shard.stats()[5])
                return [int(i) for i in idx]
            if getattr(self._comm, 'on_device', False) and not show_progress:
                if self._peer and 2 * self._shard.record_doubles() <= self._shard.peer_slot_doubles():
                    # (2 x: room for the batch's projection columns may still double the record in ital_fetch_begin)
                    idx, scores = self._shard.fetch_peer(k, self.label_prob, self.mistake_prob, self._exhaustive())
                    self.last_fetch_scores = scores
                    self.last_fused_steps = int(self._shard.stats()[5])
                    return [int(i) for i in idx]
                return self._fetch_device_loop(k)
            return self._fetch_stepwise(k, show_progress)
        finally:
            if restricted:
                self._shard.restrict_candidates(None)

    def _fetch_change_subset(self, k, keep_scores=False):                       # ital.py:102-108, 227-275, 514-584
        """Greedy loop with change_estimation_subset > 0: the reference's own draw of the subset (same call on the
        global numpy RNG), then per step one evaluation of all candidates outside ext = batch + subset and one per
        subset member that is still a candidate (scored with itself moved out of the subset)."""
        sh, comm = self._shard, self._comm
        candidates = self.get_unseen()
        S = sorted(np.random.choice(candidates, min(len(candidates), self.change_estimation_subset), replace=False))
        S = [int(i) for i in S]
        _capi.check(sh.lib.ital_set_lazy_rows(sh.handle, 0))    # every row needs its projection on ext
        sh.set_sub_mode(True)
        ret, self.last_fetch_scores, self.last_subset, self.last_step_scores = [], [], list(S), []

        def evaluate(batch, sub, only):
            sh.fetch_begin(1.0, self.mistake_prob)
            try:
                for e in batch + sub:
                    sh.fetch_commit(comm.sum_records(sh.export_points([e]))[0])
                allrec = comm.gather_records(sh.fetch_propose_sub(len(batch), only))
                self.last_fetch_stats.append(sh.stats())
                if keep_scores:
                    sc = self._all_rows(sh.last_scores())[:self._n]
                    seen = ~np.isnan(sc)
                    self.last_step_scores[-1][seen] = sc[seen]
            finally:
                sh.fetch_end()
            win = pick_winner(allrec)
            return None if win < 0 else (float(allrec[win][1]), int(allrec[win][0]))

        try:
            for it in range(k):
                sub = [i for i in S if i not in ret]
                if keep_scores:
                    self.last_step_scores.append(np.full(self._n, np.nan))
                best = evaluate(ret, sub, -1)
                for i in sub:
                    cur = evaluate(ret, [j for j in sub if j != i], i)
                    if cur is not None and (best is None or cur[0] > best[0] or (cur[0] == best[0] and cur[1] < best[1])):
                        best = cur
                if best is None:
                    break
                ret.append(best[1])
                self.last_fetch_scores.append(best[0])
        finally:
            sh.set_sub_mode(False)
        return ret

    def _fetch_device_loop(self, k):
        """Multi-GPU greedy loop with the records staying on the GPUs: per step one enqueue of the local
        scoring, one NCCL all-gather of the shards' best records, one enqueue of winner pick + streaming pass.
        The host never waits inside the loop; one read-back at the end."""
        comm = self._comm
        self._shard.fetch_begin(self.label_prob, self.mistake_prob)
        try:
            rl = self._shard.record_doubles()
            rec, allrec = comm.device_buffers(rl)
            for it in range(k):
                self._shard.fetch_propose_dev(-np.inf, self._exhaustive(), rec.data_ptr())
                comm.all_gather_device(allrec, rec)
                self._shard.fetch_commit_dev(allrec.data_ptr(), comm.world_size, it + 1 < k)
            idx, scores = self._shard.fetch_result(k)
        finally:
            self._shard.fetch_end()
        self.last_fetch_scores = scores
        return [int(i) for i in idx]

    def _fetch_stepwise(self, k, show_progress=False, keep_scores=False):
        """The greedy loop with the per-step exchange made explicit (multi-GPU, progress bars, tests)."""
        steps = range(k)
        if show_progress:
            from tqdm import trange
            steps = trange(k)
        ret, self.last_fetch_scores, self.last_step_scores = [], [], []
        self._apply_lazy_rows()
        self._apply_clip_cov()
        self._shard.fetch_begin(self.label_prob, self.mistake_prob)
        try:
            for it in steps:
                rec = self._shard.fetch_propose(-np.inf, self._exhaustive())
                self.last_fetch_stats.append(self._shard.stats())
                if keep_scores:
                    self.last_step_scores.append(self._all_rows(self._shard.last_scores())[:self._n])
                allrec = self._comm.gather_records(rec)
                win = pick_winner(allrec)
                if win < 0:
                    break
                ret.append(int(allrec[win][0]))
                self.last_fetch_scores.append(float(allrec[win][1]))
                if it + 1 < k:
                    self._shard.fetch_commit(allrec[win])
        finally:
            self._shard.fetch_end()
        return ret


class EntropySampling(ITAL):
    """Greedy maximum joint-entropy batches; drop-in for `ital.baseline_methods.EntropySampling`
    (ital/baseline_methods.py:229-287), which scores a batch by the entropy of the sign pattern of its joint GP
    posterior.  That is ITAL's mutual information under the perfect-user model, so the same kernels
    serve it; the scores differ from the reference's only in its clamps (probabilities clipped to [1e-8, 1 - 1e-8]
    for one sample, terms below 1e-12 dropped for several), i.e. by less than 1e-7."""

    def __init__(self, data=None, queries=[], length_scale=0.1, var=1.0, noise=1e-6, **kwargs):
        for name in ('label_prob', 'mistake_prob', 'label_estimation'):
            kwargs.pop(name, None)
        ITAL.__init__(self, data, queries, length_scale, var, noise, label_prob=1.0, mistake_prob=0.0, **kwargs)


class VarianceSampling(ITAL):
    """Maximum predictive variance; drop-in for `ital.baseline_methods.VarianceSampling`
    (ital/baseline_methods.py:110-155).  With `use_correlations` a batch is scored by the sum of its variances minus
    the sum of its covariances and built greedily: the change when row i joins is v_i - sum_a cov(r_a, i), read off the
    incremental Cholesky rows that ITAL's streaming pass maintains for every row (predict_cov_batch is never formed).
    Like the reference, the first pick of that mode does not exclude unnameable samples (baseline_methods.py:133)."""

    def __init__(self, data=None, queries=[], length_scale=0.1, var=1.0, noise=1e-6, use_correlations=False, **kwargs):
        self.use_correlations = use_correlations
        for name in ('label_prob', 'mistake_prob', 'label_estimation', 'lazy_rows'):
            kwargs.pop(name, None)
        ITAL.__init__(self, data, queries, length_scale, var, noise, **kwargs)

    def fetch_unlabelled(self, k, show_progress=False):
        if len(self._labelled_idx) == 0:
            raise RuntimeError('fetch_unlabelled() needs at least one query or labelled sample')
        k = int(k)
        if k > self.MAX_BATCH:
            raise NotImplementedError('batches of more than %d samples are not supported' % self.MAX_BATCH)
        corr = bool(self.use_correlations)
        # the streaming pass keeps every row's batch columns current (all rows are scored); without correlations
        # nothing of the batch enters the score, so no pass is needed at all
        _capi.check(self._shard.lib.ital_set_lazy_rows(self._shard.handle, int(not corr)))
        ret, self.last_fetch_scores = [], []
        self._shard.fetch_begin(1.0, 0.0)
        try:
            for it in range(k):
                rec = self._shard.variance_propose(corr, corr)
                allrec = self._comm.gather_records(rec)
                win = pick_winner(allrec)
                if win < 0:
                    break
                ret.append(int(allrec[win][0]))
                self.last_fetch_scores.append(float(allrec[win][1]))
                if it + 1 < k:
                    self._shard.fetch_commit(allrec[win])
        finally:
            self._shard.fetch_end()
        return ret
