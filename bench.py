"""Benchmark of the ITAL batch-selection path: `fetch_unlabelled(4)` on the synthetic SYN pool
(n = 1M rows per GPU, d = 512, float32-representable, |L| = 9 labelled, perfect-user model).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--total-rows R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one `fetch_unlabelled(batch)` call on the prepared learner (the call does not change the model, so
every step does the same work).  `value` = candidates ranked per second = sum over greedy steps of the unseen
candidates the step had to rank / device-timed latency (CUDA events around every call, L2 flushed between calls,
max over ranks), whole job.  The product ranks them with an exact lazy-greedy bound (same batch as scoring every
candidate), in one persistent kernel per fetch; the JSON also carries the same metric for the other
modes of the same call: `exhaustive` (every candidate scored by quadrature each step -- what the CPU arms do),
`streaming` (one HBM pass over the pool per greedy step, the round-1 default) and `multi_kernel` (the fused
kernel's phases as separate launches).  `e2e` is wall clock through the public `ITAL.fetch_unlabelled` (host
arguments in, host list out, every host<->device copy of the call inside the timed region, bytes counted by the
library).  `roofline` rates the HBM-bound streaming pass where its output is consumed -- the single-column pass of
the exhaustive fetch, whose new projection entries every candidate's score reads -- by its algorithmic bytes over its
own CUDA-event time; `roofline_update` does the same for the multi-column labelled pass of `update` (consumed every
active-learning round) and `roofline_streaming` for the pass inside the pipelined streaming fetch.  `cpu_baseline` / `--impl reference` time the float64 oracle
(the port of the reference's algorithm; the reference itself cannot allocate its n-by-n kernel matrix at this size
and is Python that does not travel to the GPU box) on a bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'MI candidate-scores/sec (fetch_unlabelled, n=1M rows per GPU, d=512, batch=4)'
UNIT = 'candidates/s'
# fetch_unlabelled(4) of the default workload on ONE GPU (1M rows, |L| = 9): what every sharding of the same pool
# must return (strong scaling: --total-rows 1000000)
BATCH_1M = [956886, 394323, 849682, 347341]


def syn_block(first_row, rows, d, centres=1000, seed=0):
    """Rows [first_row, first_row + rows) of the SYN pool: clustered, L2-normalised, float32.

    Centres come from `seed`; every block of 65536 rows has its own stream so that any shard of the pool can
    be generated independently and reproducibly."""
    C = np.random.default_rng(seed).standard_normal((centres, d)).astype(np.float32)
    out = np.empty((rows, d), dtype=np.float32)
    assign = np.empty(rows, dtype=np.int64)
    blk = 65536
    b = first_row // blk
    pos = 0
    while pos < rows:
        g0 = max(first_row, b * blk)
        g1 = min(first_row + rows, (b + 1) * blk)
        rng = np.random.default_rng([seed, 1, b])
        a = rng.integers(0, centres, blk)
        noise = rng.standard_normal((blk, d), dtype=np.float32)
        sl = slice(g0 - b * blk, g1 - b * blk)
        x = C[a[sl]] + np.float32(0.6) * noise[sl]
        x /= np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
        out[pos:pos + (g1 - g0)] = x
        assign[pos:pos + (g1 - g0)] = a[sl]
        pos += g1 - g0
        b += 1
    return out, assign


def labelled_state(assign_head):
    """|L| = 9: the query (row 0), four more rows of its cluster (+1), the first rows of four other clusters
    (-1); all taken from the head of the pool."""
    c0 = assign_head[0]
    pos = [int(i) for i in np.nonzero(assign_head == c0)[0][1:5]]
    neg, seen = [], {c0}
    for i, c in enumerate(assign_head):
        if c not in seen:
            seen.add(c)
            neg.append(int(i))
        if len(neg) == 4:
            break
    return [{0: 1}, {**{i: 1 for i in pos}, **{i: -1 for i in neg}}]


def candidates_ranked(n_total, n_labelled, batch):
    return sum(n_total - n_labelled - t for t in range(batch))


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        fd, self.path = tempfile.mkstemp(suffix='.csv')
        self.out = os.fdopen(fd, 'w')
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            self.out.close()
            os.unlink(self.path)
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.out.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def ncu_traffic(pattern):
    """DRAM bytes per launch from the newest committed `ncu --set full` summary matching `pattern`, or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', pattern)))
    if not files:
        return None, None
    rd = wr = None
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    for line in open(files[-1]):
        f = line.split()
        try:
            if line.startswith('dram__bytes_read.sum') and f[1] in scale:
                rd = np.mean([float(x) for x in f[2:]]) * scale[f[1]]
            if line.startswith('dram__bytes_write.sum') and f[1] in scale:
                wr = np.mean([float(x) for x in f[2:]]) * scale[f[1]]
        except (ValueError, IndexError):
            continue
    if rd is None or wr is None:
        return None, None
    return float(rd + wr), os.path.relpath(files[-1], ROOT)


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        return 6650.0, 'fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)'


def _one_blas_thread():
    """Pool workers run one BLAS thread each (README.md:78-87 of the reference: MKL/OMP/OPENBLAS_NUM_THREADS=1 when
    parallelized=True); the limiter object has to stay alive for the life of the worker."""
    global _BLAS_LIMIT
    from threadpoolctl import threadpool_limits
    _BLAS_LIMIT = threadpool_limits(limits=1)


def oracle_fetch_rate(rows, d, batch, procs, steps=1, warmup=0, budget_s=None):
    """Candidates/s of the float64 oracle on the first `rows` rows of the same pool, `procs` host processes."""
    import multiprocessing as mp
    from oracle.ital_oracle import OracleITAL
    X, assign = syn_block(0, rows, d)
    ora = OracleITAL(X.astype(np.float64), length_scale=1.0)
    for fb in labelled_state(assign[:65536]):
        ora.update(fb)
    pool = mp.get_context('fork').Pool(procs, initializer=_one_blas_thread) if procs > 1 else None
    try:
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            ret = ora.fetch_unlabelled(batch, pool=pool)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
            if budget_s is not None and sum(times) > budget_s and times:
                break
    finally:
        if pool is not None:
            pool.close()
    per = float(np.mean(times))
    return candidates_ranked(rows, 9, batch) / per, per, ret, len(times)


def oracle_rescore(X_rows, fbs_local, batch_local):
    """Scores of the selected batch along its own greedy path, by the float64 oracle on the handful of rows involved
    (labelled rows + batch): the score of ret[:t+1] depends on those rows only."""
    from oracle.ital_oracle import OracleITAL
    ora = OracleITAL(X_rows.astype(np.float64), length_scale=1.0)
    for fb in fbs_local:
        ora.update(fb)
    ora.fetch_unlabelled(len(batch_local), forced=list(batch_local))
    out = []
    for t, tr in enumerate(ora.trace):
        pos = int(np.nonzero(tr['candidates'] == batch_local[t])[0][0])
        out.append(float(tr['scores'][pos]))
    return out


def run_reference(args, rank, world):
    """The reference arm: the oracle port of the reference's CPU algorithm on all host cores, bounded sample."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = args.cpu_rows if args.cpu_rows else 16000
    rate, per, ret, done = oracle_fetch_rate(rows, args.dim, args.batch, cores, steps=args.steps,
                                             warmup=min(args.warmup, 1), budget_s=150.0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': done,
        'warmup': min(args.warmup, 1), 'ms_per_step': per * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'SYN pool, fetch_unlabelled(%d), |L|=9, perfect user; bounded sample: first %d rows '
                               'of the 1M-row pool, every candidate scored each step' % (args.batch, rows),
                   'rows': rows, 'd': args.dim, 'batch': args.batch},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'first %d rows of SYN-1M, %d timed fetches; numpy float64 oracle, candidates '
                                   'spread over a fork pool of %d processes (the reference itself is Python, '
                                   'needs an n-by-n matrix and the removed scipy mvndst: it cannot run this '
                                   'size nor travel to the GPU box)' % (rows, done, cores)},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


class Harness(object):
    """One prepared learner on this rank's shard plus the timing helpers."""

    def __init__(self, args, rank, world, local_rank, rows_per_gpu, group, **learner_kw):
        import torch
        from ital_b200 import ITAL
        self.torch, self.args, self.rank, self.world = torch, args, rank, world
        self.rows = rows_per_gpu
        self.n_total = rows_per_gpu * world
        self.first = rank * rows_per_gpu
        t0 = time.perf_counter()
        self.X, self.assign = syn_block(self.first, rows_per_gpu, args.dim)
        # labels come from the pool's head (rank 0 owns it when shards hold >= 65536 rows; otherwise regenerate)
        self.head = self.assign[:65536] if rank == 0 and rows_per_gpu >= 65536 else syn_block(0, 65536, args.dim)[1]
        self.t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.learner = ITAL(self.X, length_scale=1.0, device=local_rank, process_group=group,
                            local_rows=(self.first, self.n_total) if world > 1 else None, **learner_kw)
        self.t_fit = time.perf_counter() - t0
        self.fbs = labelled_state(self.head)
        t0 = time.perf_counter()
        for fb in self.fbs:
            self.learner.update(fb)
        torch.cuda.synchronize()
        self.t_update = time.perf_counter() - t0
        self.n_lab = sum(len(f) for f in self.fbs)
        self.shard = self.learner._shard
        self.lib = self.shard.lib
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')   # 2 x the 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def set_mode(self, exhaustive=False, lazy_rows=None, fused=True):
        self.learner.exhaustive = exhaustive
        self.learner.lazy_rows = lazy_rows
        self.learner.fused = fused

    def timed(self, steps, batch=None, flush=True):
        """`steps` fetches: device time by a CUDA-event pair around every call on the stream the kernels use (L2
        flushed before each call, outside the pair), wall clock of every call beside it; max over ranks."""
        torch = self.torch
        batch = self.args.batch if batch is None else batch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        wall = 0.0
        ret = None
        for e0, e1 in evs:
            if flush:
                self.flush_buf.fill_(1)           # 256 MB written: nothing of the previous call is left in L2
                torch.cuda.synchronize()
            w0 = time.perf_counter()
            e0.record()
            ret = self.learner.fetch_unlabelled(batch)
            e1.record()
            torch.cuda.synchronize()
            wall += time.perf_counter() - w0
        dev = sum(e0.elapsed_time(e1) for e0, e1 in evs) * 1e-3
        if self.world > 1:
            import torch.distributed as dist
            tt = torch.tensor([dev, wall], device='cuda', dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dev, wall = float(tt[0]), float(tt[1])
            dist.barrier()
        return dev, wall, ret

    def transfer_bytes(self):
        h, d = ctypes.c_int64(), ctypes.c_int64()
        self.lib.ital_transfer_bytes(self.shard.handle, ctypes.byref(h), ctypes.byref(d))
        return h.value, d.value

    def profile_read(self):
        ms, nl, nb = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
        self.lib.ital_profile_read(self.shard.handle, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(nb))
        return ms.value, nl.value, nb.value

    def gather_rows(self, idx):
        """Rows of the pool by global index, on every rank (a handful of rows for the oracle re-score)."""
        out = np.zeros((len(idx), self.args.dim), dtype=np.float32)
        for k, i in enumerate(idx):
            if self.first <= i < self.first + self.rows:
                out[k] = self.X[i - self.first]
        if self.world > 1:
            import torch.distributed as dist
            t = self.torch.from_numpy(out).cuda()
            dist.all_reduce(t)
            out = t.cpu().numpy()
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rows', type=int, default=1000000, help='pool rows per GPU (weak scaling)')
    ap.add_argument('--total-rows', type=int, default=0, help='pool rows in total, split over the GPUs (strong scaling)')
    ap.add_argument('--dim', type=int, default=512)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--cpu-rows', type=int, default=0, help='rows of the CPU baseline sample (0 = default)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--exhaustive-steps', type=int, default=1)
    ap.add_argument('--secondary-steps', type=int, default=30, help='timed fetches of the streaming / multi-kernel rows')
    ap.add_argument('--rounds', type=int, default=8, help='full fetch+update rounds timed at the end (8 rounds of 4: |L| = 41)')
    ap.add_argument('--model-steps', type=int, default=20, help='timed fetches with mistake_prob = 0.5')
    ap.add_argument('--batch10-steps', type=int, default=1, help='timed fetches of a batch of 10')
    ap.add_argument('--general-steps', type=int, default=1, help='timed fetches with label_prob = 0.25 (slow: every candidate scored)')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling sub-run at N > 1')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    group = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        group = True

    strong = args.total_rows > 0
    rows = args.total_rows // world if strong else args.rows
    H = Harness(args, rank, world, local_rank, rows, group)
    learner, shard, lib = H.learner, H.shard, H.lib
    n_total, n_lab = H.n_total, H.n_lab
    ranked = candidates_ranked(n_total, n_lab, args.batch)

    # clocks are sampled from the warm-up on: the timed region itself can be shorter than one nvidia-smi period
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    # ---- headline: the default mode (projections on demand, one persistent kernel per fetch) -----------------------
    H.set_mode()
    H.timed(args.warmup)                                    # warm-up (also grows every buffer to its steady size)
    launches0 = int(lib.ital_launch_count(shard.handle))
    h2d0, d2h0 = H.transfer_bytes()
    dev, wall, ret = H.timed(args.steps)
    clocks = sampler.stop() if sampler else None
    launches = int(lib.ital_launch_count(shard.handle)) - launches0
    h2d1, d2h1 = H.transfer_bytes()
    h2d, d2h = (h2d1 - h2d0) // args.steps, (d2h1 - d2h0) // args.steps
    fused_steps = int(learner.last_fused_steps)
    scores = [float(x) for x in learner.last_fetch_scores]
    warm_dev, warm_wall, _ = H.timed(min(args.steps, 50), flush=False)      # back to back, vectors resident in L2
    stats = None
    if world == 1:
        learner._fetch_stepwise(args.batch)
        stats = [[float(x) for x in s[:7]] for s in learner.last_fetch_stats]

    # ---- correctness carried by the line ---------------------------------------------------------------------------
    checks = {}
    lab_idx = [i for fb in H.fbs for i in fb]
    rows_needed = lab_idx + [int(i) for i in ret]
    Xs = H.gather_rows(rows_needed)
    if rank == 0:
        pos = {g: k for k, g in enumerate(rows_needed)}
        want = oracle_rescore(Xs, [{pos[i]: v for i, v in fb.items()} for fb in H.fbs], [pos[int(i)] for i in ret])
        err = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(scores, want))
        checks['batch_scores_vs_oracle_max_rel_err'] = err
        checks['batch_scores_match_oracle_1e-6'] = bool(err <= 1e-6)
    if n_total == 1000000 and args.dim == 512 and args.batch == 4:
        checks['batch_matches_n1'] = [int(i) for i in ret] == BATCH_1M

    # ---- the other modes of the same call ---------------------------------------------------------------------------
    def row(dev_s, wall_s, steps, r, note):
        return {'value': ranked * steps / dev_s, 'unit': UNIT, 'ms_per_step': dev_s / steps * 1e3,
                'e2e_ms_per_step': wall_s / steps * 1e3, 'steps': steps, 'same_batch': r == ret, 'note': note}

    peak, peak_src = measured_peak()
    multi = streaming = exh = roof_stream = None
    if args.secondary_steps > 0:
        H.set_mode(lazy_rows=True, fused=False)
        H.timed(3)
        mdev, mwall, mret = H.timed(args.secondary_steps)
        multi = row(mdev, mwall, args.secondary_steps, mret,
                    'ITAL_B200_FUSED=0: the same phases as ~30 dependent kernel launches per fetch (round-1 lazy_rows)')
        multi['scores_bit_identical'] = [float(x) for x in learner.last_fetch_scores] == scores
        H.set_mode(lazy_rows=False)
        H.timed(3)
        lib.ital_profile_enable(shard.handle, 1)
        H.profile_read()
        sdev, swall, sret = H.timed(args.secondary_steps)
        sms, snl, snb = H.profile_read()
        lib.ital_profile_enable(shard.handle, 0)
        streaming = row(sdev, swall, args.secondary_steps, sret,
                        'lazy_rows=False: one HBM pass over the pool per greedy step keeps every row\'s projection '
                        'current (k_extend_bulk, pipelined with the scoring chain); what exhaustive scoring and the '
                        'general feedback model consume, and the round-1 default')
        streaming['scores_bit_identical'] = [float(x) for x in learner.last_fetch_scores] == scores
        s_ach = (snb / 1e9) / (sms / 1e3) if sms > 0 else 0.0
        s_traffic, s_src = ncu_traffic('r*_k_extend_bulk_full.txt')
        roof_stream = {'bound': 'hbm', 'kernel': 'k_extend_bulk<float,4> (single-column streaming pass of a fetch)',
                       'achieved': s_ach, 'peak': peak, 'unit': 'GB/s', 'frac': s_ach / peak, 'traffic': s_traffic,
                       'traffic_source': s_src, 'launches': int(snl), 'avg_launch_ms': sms / max(1, snl),
                       'algorithmic_bytes_per_launch': snb / max(1, snl), 'share_of_step': sms / 1e3 / sdev}
    roof_consumed = None
    if args.exhaustive_steps > 0:
        H.set_mode(exhaustive=True)
        lib.ital_profile_enable(shard.handle, 1)
        H.profile_read()
        edev, ewall, eret = H.timed(args.exhaustive_steps)
        ems, enl, enb = H.profile_read()
        lib.ital_profile_enable(shard.handle, 0)
        exh = row(edev, ewall, args.exhaustive_steps, eret,
                  'every candidate scored by quadrature at every greedy step (no lazy-greedy bound): the like-for-like '
                  'count against the CPU arms; one streaming pass per step keeps all projections current')
        checks['exhaustive_same_batch'] = eret == ret
        if stats is not None:
            # FP64-pipe roofline of the scorer (the MI arithmetic is the secondary, compute bound): node
            # evaluations per second x FP64-pipe instructions per node (25, counted in the SASS of k_eval<3>:
            # 3 DFMA for the argument, 1 DMUL, ~20 for the table-based normal CDF, 1 DFMA accumulate) against the DFMA
            # issue rate measured on this GPU (tools/probe/dmma_bench.cu: 0.55 cycles per warp instruction and SM)
            node_evals = sum((n_total - n_lab - t) * pad for t, pad in enumerate([0] + [s[6] for s in stats[1:]]))
            clk = (clocks or {}).get('sm_mhz') or 1965.0
            peak_fp64 = 148 * 32 / 0.55 * clk * 1e6
            ach = node_evals * 25 * args.exhaustive_steps / edev
            exh['fp64_roofline'] = {'bound': 'fp64 pipe', 'node_evaluations_per_fetch': node_evals,
                                    'fp64_instructions_per_node': 25, 'achieved': ach, 'peak': peak_fp64,
                                    'unit': 'FP64 lane-instructions/s', 'frac': ach / peak_fp64,
                                    'ncu': 'profiles/r02_k_eval_full.txt: sm__pipe_fp64_cycles_active 51 %, issue slots 49 %, XU 16 %'}
        e_ach = (enb / 1e9) / (ems / 1e3) if ems > 0 else 0.0
        e_traffic, e_src = ncu_traffic('r*_k_extend_bulk_full.txt')
        roof_consumed = {'bound': 'hbm',
                         'kernel': 'k_extend_bulk<float,4> -- the single-column streaming pass, measured where its output is '
                                   'consumed: exhaustive scoring reads every row\'s new projection entry in the next step '
                                   '(so does the general feedback model, label_prob < 1)',
                         'achieved': e_ach, 'peak': peak, 'unit': 'GB/s', 'frac': e_ach / peak, 'traffic': e_traffic,
                         'traffic_source': e_src, 'peak_source': peak_src, 'launches': int(enl),
                         'avg_launch_ms': ems / max(1, enl), 'algorithmic_bytes_per_launch': enb / max(1, enl),
                         'timed_region': 'the %d exhaustive fetch(es) of this run (CUDA events inside the library)' % args.exhaustive_steps,
                         'share_of_step': ems / 1e3 / edev}
    H.set_mode()

    # ---- other feedback models on the same pool (SURVEY.md 8d secondary rows) ---------------------------------------
    models = {}
    if args.model_steps > 0:
        learner.mistake_prob = 0.5                      # configs/butterflies-aggressive.conf: labels wrong half of the time
        H.timed(2)
        mdev2, mwall2, mret2 = H.timed(args.model_steps)
        models['mistake_prob_0.5'] = row(mdev2, mwall2, args.model_steps, mret2,
                                         'label_prob=1, mistake_prob=0.5: perfect-user scores + a per-step constant (same '
                                         'argmax), persistent kernel')
        learner.mistake_prob = 0.0
    if args.batch10_steps > 0:
        r10 = candidates_ranked(n_total, n_lab, 10)
        H.timed(1, batch=10, flush=False)
        bdev, bwall, bret = H.timed(args.batch10_steps, batch=10)
        models['batch_10'] = {'value': r10 * args.batch10_steps / bdev, 'unit': UNIT, 'ms_per_step': bdev / args.batch10_steps * 1e3,
                              'e2e_ms_per_step': bwall / args.batch10_steps * 1e3, 'steps': args.batch10_steps,
                              'batch': [int(i) for i in bret], 'first_four_match': [int(i) for i in bret[:4]] == [int(i) for i in ret],
                              'note': 'fetch_unlabelled(10) (BASELINE.json config 5): the persistent kernel runs the first four '
                                      'steps; steps 5-10 run the multi-kernel loop with node sets generated on the device (tensor '
                                      'rule with 78k / 386k kept nodes for 4 / 5 base variables, sequential-conditioning lattice '
                                      'with 2^18 nodes from 6 on) -- the time is the exact scoring of the 2 500 - 11 000 rows per '
                                      'step that the wider pruning margin of those rules (2e-5 / 1e-3) lets through'}
    if args.general_steps > 0:
        learner.label_prob = 0.25                       # configs/butterflies-conservative.conf: every candidate is scored
        gdev, gwall, gret = H.timed(args.general_steps, batch=min(args.batch, 4), flush=False)
        models['label_prob_0.25'] = row(gdev, gwall, args.general_steps, gret,
                                        'label_prob=0.25 (general feedback model): conditional node sets (~20.7k nodes at '
                                        'the 4th step), every candidate scored at every step, streaming passes')
        models['label_prob_0.25']['same_batch'] = None
        models['label_prob_0.25']['batch'] = [int(i) for i in gret]
        learner.label_prob = 1.0

    # ---- a few full active-learning rounds (fetch 4, label them, update): the update is the consumer of the
    # streaming pass (one multi-column pass per <= 4 labels); this changes the model, so it runs last ----------------
    rounds = roof = None
    if args.rounds > 0:
        c0 = H.head[0]
        H.barrier()
        upd_ms, upd_gbs, upd_wall, fetch_each = [], [], [], []
        tot_ms = tot_bytes = 0.0
        tot_launch = 0
        for _ in range(args.rounds):
            w0 = time.perf_counter()
            batch = learner.fetch_unlabelled(args.batch)
            w1 = time.perf_counter()
            lab = {}
            for i in batch:        # simulated user: relevance = membership in the query's cluster
                owner = i // rows
                ci = H.assign[i - H.first] if owner == rank else -1
                if world > 1:
                    tt = torch.tensor([ci], device='cuda')
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    ci = int(tt[0])
                lab[i] = 1 if ci == c0 else -1
            lib.ital_profile_enable(shard.handle, 1)
            H.profile_read()
            torch.cuda.synchronize()
            w2 = time.perf_counter()
            learner.update(lab)
            torch.cuda.synchronize()
            w3 = time.perf_counter()
            ums, unl, unb = H.profile_read()
            lib.ital_profile_enable(shard.handle, 0)
            tot_ms += ums
            tot_bytes += unb
            tot_launch += unl
            upd_ms.append(ums / max(1, unl))
            upd_gbs.append(unb / 1e9 / (ums / 1e3) if ums > 0 else 0.0)
            upd_wall.append((w3 - w2) * 1e3)
            fetch_each.append((w1 - w0) * 1e3)
        learner.top_results(100)                      # first call allocates the sort buffers
        tops = []
        for _ in range(10):
            w0 = time.perf_counter()
            learner.top_results(100)
            tops.append(time.perf_counter() - w0)
        rounds = {'rounds': args.rounds, 'fetch_ms': float(np.median(fetch_each)), 'fetch_ms_max': float(np.max(fetch_each)),
                  'update_ms': float(np.median(upd_wall)), 'update_ms_each': upd_wall,
                  'top_results_100_ms': float(np.median(tops)) * 1e3,
                  'fetch_ms_each': fetch_each,
                  'update_pass_ms': float(np.median(upd_ms)), 'update_pass_GBs': float(np.median(upd_gbs)),
                  'labelled_after': n_lab + args.rounds * args.batch,
                  'note': 'wall clock; fetch right after an update, i.e. with a cold L2 (median; the maximum contains the '
                          'one-off growth of the projection matrix when the column capacity doubles); update(%d labels) = '
                          'one multi-column streaming pass + bookkeeping; top_results(100) = device radix sort of the '
                          'local means, 100 indices read back' % args.batch}
        H.set_mode()
        H.timed(2)
        ldev, lwall, lret = H.timed(min(args.steps, 50))
        r41 = candidates_ranked(n_total, n_lab + args.rounds * args.batch, args.batch)
        rounds['fetch_after_rounds'] = {'labelled': n_lab + args.rounds * args.batch,
                                        'value': r41 * min(args.steps, 50) / ldev, 'unit': UNIT,
                                        'ms_per_step': ldev / min(args.steps, 50) * 1e3,
                                        'e2e_ms_per_step': lwall / min(args.steps, 50) * 1e3,
                                        'batch': [int(i) for i in lret],
                                        'note': 'the headline measurement repeated on the model after the rounds (|L| = %d)'
                                                % (n_lab + args.rounds * args.batch)}
        u_ach = (tot_bytes / 1e9) / (tot_ms / 1e3) if tot_ms > 0 else 0.0
        u_traffic, u_src = ncu_traffic('r*_k_extend_bulk_multi_full.txt')
        roof = {'bound': 'hbm',
                'kernel': 'k_extend_bulk_multi<float,4,%d> -- the labelled pass of update(): the multi-column HBM pass that '
                          'is consumed every active-learning round' % args.batch,
                'achieved': u_ach, 'peak': peak, 'unit': 'GB/s', 'frac': u_ach / peak, 'traffic': u_traffic,
                'traffic_source': u_src, 'peak_source': peak_src, 'launches': int(tot_launch),
                'avg_launch_ms': tot_ms / max(1, tot_launch),
                'algorithmic_bytes_per_launch': tot_bytes / max(1, tot_launch),
                'timed_region': 'the %d update() calls of al_rounds in this run (CUDA events inside the library)' % args.rounds}
    roof_update = roof
    roof = roof_consumed or roof_update or roof_stream

    used_peer = bool(getattr(learner, '_peer', False))
    t_fit, t_gen, t_update, x_bytes = H.t_fit, H.t_gen, H.t_update, H.X.nbytes
    learner.close()
    del H

    # ---- strong scaling beside the weak-scaling line: the 1M-row pool of the metric split over the GPUs ------------
    strong_row = None
    if world > 1 and not strong and not args.no_strong and args.rows == 1000000:
        HS = Harness(args, rank, world, local_rank, 1000000 // world, group)
        HS.set_mode()
        HS.timed(args.warmup)
        sdev2, swall2, sret2 = HS.timed(args.steps)
        r1 = candidates_ranked(1000000 // world * world, n_lab, args.batch)
        strong_row = {'total_rows': 1000000 // world * world, 'rows_per_gpu': 1000000 // world,
                      'value': r1 * args.steps / sdev2, 'unit': UNIT, 'ms_per_step': sdev2 / args.steps * 1e3,
                      'e2e_ms_per_step': swall2 / args.steps * 1e3, 'steps': args.steps,
                      'batch': [int(i) for i in sret2],
                      'batch_matches_n1': [int(i) for i in sret2] == BATCH_1M if 1000000 % world == 0 else None,
                      'note': 'the n = 1M pool of the metric sharded over %d GPUs (strong scaling); every sharding must '
                              'return the 1-GPU batch' % world}
        HS.learner.close()
        del HS

    if rank != 0:
        dist.destroy_process_group()
        return
    line = {
        'metric': METRIC, 'value': ranked * args.steps / dev, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dev / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'SYN-1M: %d rows per GPU x %d GPU(s), d=%d float32-representable features in HBM, '
                               'fetch_unlabelled(%d), |L|=%d, length_scale=1, perfect user; float64 arithmetic'
                               % (rows, world, args.dim, args.batch, n_lab),
                   'rows_per_gpu': rows, 'total_rows': n_total, 'd': args.dim, 'batch': args.batch, 'labelled': n_lab,
                   'candidates_ranked_per_step': ranked,
                   'counting': 'value = candidates ranked exactly per second (lazy-greedy bound: a few hundred of them '
                               'need the quadrature); `exhaustive` counts candidates scored by quadrature, as the CPU arms do',
                   'l2': 'L2 flushed (256 MB written) before every timed fetch: the default fetch reads per-row vectors '
                         '(25 MB) that would otherwise stay in the 126 MB L2; warm_l2_ms_per_step is the back-to-back figure',
                   'mode': 'default: projections on demand + fused persistent kernel (%d of %d greedy steps fused)'
                           % (fused_steps, args.batch),
                   'parallelism': 'rows sharded over %d GPU(s); per greedy step every shard %s' % (
                       world, 'stores its proposal into its peers\' memory over NVLink from inside the persistent kernel '
                              '(CUDA IPC; no NCCL in the loop)' if used_peer else
                       ('needs no exchange (one shard)' if world == 1 else 'contributes one record to an NCCL all-gather')),
                   'batch_selected': [int(i) for i in ret], 'batch_scores': scores},
        'e2e': {'value': ranked * args.steps / wall, 'unit': UNIT, 'ms_per_step': wall / args.steps * 1e3,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'note': 'wall clock of ITAL.fetch_unlabelled through the Python/ctypes/C-ABI boundary, bytes counted by '
                        'the library (ital_transfer_bytes); the pool itself is resident (uploaded once by fit: %.0f ms, '
                        '%.2f GB)' % (t_fit * 1e3, x_bytes / 1e9)},
        'warm_l2_ms_per_step': warm_dev / min(args.steps, 50) * 1e3,
        'warm_l2_e2e_ms_per_step': warm_wall / min(args.steps, 50) * 1e3,
        'gpu_launches': launches,
        'clocks': clocks,
        'checks': checks,
        'roofline': roof,
        'roofline_update': roof_update,
        'roofline_streaming': roof_stream,
        'exhaustive': exh,
        'streaming': streaming,
        'multi_kernel': multi,
        'feedback_models': models,
        'al_rounds': rounds,
        'strong_scaling': strong_row,
        'fetch_stats_per_step': stats,
        'setup': {'generate_s': t_gen, 'fit_s': t_fit, 'update_9_labels_s': t_update},
    }
    if world == 1 and not args.no_cpu_baseline:
        crows = args.cpu_rows if args.cpu_rows else 12000
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            rate, per, cret, done = oracle_fetch_rate(crows, args.dim, args.batch, 1, steps=1)
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                                'sample': 'first %d rows of the same pool, one fetch_unlabelled(%d) with every '
                                          'candidate scored (%.1f s); numpy float64 oracle, single process'
                                          % (crows, args.batch, per)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
