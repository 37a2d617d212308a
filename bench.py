"""Benchmark of the ITAL batch-selection path: `fetch_unlabelled(4)` on the synthetic SYN pool
(n = 1M rows per GPU, d = 512, float32-representable, |L| = 9 labelled, perfect-user model).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one `fetch_unlabelled(batch)` call on the prepared learner (the call does not change the model, so
every step does the same work).  `value` = candidates ranked per second = sum over greedy steps of the
unseen candidates the step had to rank / device-timed latency (CUDA events, max over ranks), whole job.
The product ranks them with an exact lazy-greedy bound (same batch as scoring every candidate),
so the JSON also carries `exhaustive`: the same metric with every candidate scored by quadrature each step.
`e2e` is wall-clock through the public `ITAL.fetch_unlabelled` (host arguments in, host list out, every
host<->device copy of the call inside the timed region).  `roofline` rates the dominant kernel (the streaming
pass `k_extend`) by its algorithmic bytes over its own CUDA-event time.  `cpu_baseline` / `--impl reference`
time the float64 oracle (the port of the reference's algorithm; the reference itself cannot allocate its
n-by-n kernel matrix at this size and is Python that does not travel to the GPU box) on a bounded sample.
"""
import argparse
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
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
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


def ncu_traffic():
    """DRAM bytes per k_extend launch from the committed `ncu --set full` capture (profiles/), or None."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_k_extend_bulk_full.txt'))) or \
        sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_k_extend*_full.txt')))
    if not files:
        return None, None
    rd = wr = None
    for line in open(files[-1]):
        f = line.split()
        if line.startswith('dram__bytes_read.sum'):
            rd = np.mean([float(x) for x in f[2:]]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[f[1]]
        if line.startswith('dram__bytes_write.sum'):
            wr = np.mean([float(x) for x in f[2:]]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[f[1]]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rows', type=int, default=1000000, help='pool rows per GPU')
    ap.add_argument('--dim', type=int, default=512)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--cpu-rows', type=int, default=0, help='rows of the CPU baseline sample (0 = default)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--exhaustive-steps', type=int, default=1)
    ap.add_argument('--lazy-steps', type=int, default=50)
    ap.add_argument('--rounds', type=int, default=5, help='full fetch+update rounds timed at the end')
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        group = True
    from ital_b200 import ITAL

    n_total = args.rows * world
    first = rank * args.rows
    t0 = time.perf_counter()
    X, assign = syn_block(first, args.rows, args.dim)
    head = assign[:65536] if rank == 0 else syn_block(0, 65536, 8)[1]    # labels come from the pool's head
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    learner = ITAL(X, length_scale=1.0, device=local_rank, process_group=group,
                   local_rows=(first, n_total) if world > 1 else None)
    t_fit = time.perf_counter() - t0
    fbs = labelled_state(head)
    t0 = time.perf_counter()
    for fb in fbs:
        learner.update(fb)
    torch.cuda.synchronize()
    t_update = time.perf_counter() - t0
    n_lab = sum(len(f) for f in fbs)
    shard = learner._shard
    lib = shard.lib

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(steps, exhaustive):
        """`steps` fetches: device time by CUDA events on the stream the kernels use, wall clock beside it."""
        learner.exhaustive = exhaustive
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        ev0.record()
        ret = None
        for _ in range(steps):
            ret = learner.fetch_unlabelled(args.batch)
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        dev = ev0.elapsed_time(ev1) * 1e-3
        if world > 1:
            tt = torch.tensor([dev, wall], device='cuda', dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dev, wall = float(tt[0]), float(tt[1])
            dist.barrier()
        return dev, wall, ret

    # clocks are sampled from the warm-up on: the timed region itself can be shorter than one nvidia-smi period
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    # warm-up (also grows every buffer to its steady-state size)
    timed(args.warmup, False)
    # X per GPU (2 GB at 1M x 512 x 4 B) is far larger than the 126 MB L2: no L2 flush needed between steps
    import ctypes
    lib.ital_profile_enable(shard.handle, 1)
    lib.ital_profile_read(shard.handle, None, None, None)
    launches0 = int(lib.ital_launch_count(shard.handle))
    dev, wall, ret = timed(args.steps, False)
    clocks = sampler.stop() if sampler else None
    launches = int(lib.ital_launch_count(shard.handle)) - launches0
    ms = ctypes.c_double()
    nl = ctypes.c_int64()
    nbytes = ctypes.c_double()
    lib.ital_profile_read(shard.handle, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(nbytes))
    lib.ital_profile_enable(shard.handle, 0)
    stats = None
    if world == 1:
        learner._fetch_stepwise(args.batch)
        stats = [[float(x) for x in s[:5]] for s in learner.last_fetch_stats]

    ranked = candidates_ranked(n_total, n_lab, args.batch)
    value = ranked * args.steps / dev
    peak, peak_src = measured_peak()
    achieved = (nbytes.value / 1e9) / (ms.value / 1e3) if ms.value > 0 else 0.0
    traffic, traffic_src = ncu_traffic()
    # host<->device bytes of one fetch through the public API (counted from the buffers the library copies): the
    # greedy loop is device-resident (nodes, winner records and the batch state never leave the GPU), so a call
    # uploads 16 bytes of step-0 constants and reads back the selection list and the per-step counters.  With more
    # than one GPU the records cross NVLink inside the NCCL all-gather, not the host.
    h2d = 16
    d2h = 32 * 8 + 16 * 4 * 4

    exh = None
    if args.exhaustive_steps > 0:
        edev, ewall, eret = timed(args.exhaustive_steps, True)
        exh = {'value': ranked * args.exhaustive_steps / edev, 'unit': UNIT, 'ms_per_step': edev / args.exhaustive_steps * 1e3,
               'steps': args.exhaustive_steps, 'same_batch': eret == ret,
               'note': 'every candidate scored by quadrature at every greedy step (no lazy-greedy bound)'}
        learner.exhaustive = False

    lazy = None
    if args.lazy_steps > 0:
        learner.lazy_rows = True
        timed(3, False)
        ldev, lwall, lret = timed(args.lazy_steps, False)
        lazy = {'value': ranked * args.lazy_steps / ldev, 'unit': UNIT, 'ms_per_step': ldev / args.lazy_steps * 1e3,
                'e2e_ms_per_step': lwall / args.lazy_steps * 1e3, 'steps': args.lazy_steps, 'same_batch': lret == ret,
                'note': 'lazy_rows=True: batch-conditional projections extended on demand for the scored rows only '
                        '(k_catchup) instead of one streaming pass over the pool per greedy step (k_extend)'}
        learner.lazy_rows = False

    # a few full active-learning rounds (fetch 4, label them, update) for orientation: the update is the other
    # user of the streaming pass (one pass per <= 4 labels); this changes the model, so it runs last
    rounds = None
    if args.rounds > 0:
        c0 = head[0]
        barrier()
        tf = tu = tt_top = 0.0
        upd_ms, upd_gbs, fetch_each = [], [], []
        for _ in range(args.rounds):
            w0 = time.perf_counter()
            batch = learner.fetch_unlabelled(args.batch)
            w1 = time.perf_counter()
            lab = {}
            for i in batch:        # simulated user: relevance = membership in the query's cluster
                owner = i // args.rows
                ci = assign[i - first] if owner == rank else -1
                if world > 1:
                    tt = torch.tensor([ci], device='cuda')
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    ci = int(tt[0])
                lab[i] = 1 if ci == c0 else -1
            lib.ital_profile_enable(shard.handle, 1)
            learner.update(lab)
            torch.cuda.synchronize()
            w2 = time.perf_counter()
            ums, unl, unb = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
            lib.ital_profile_read(shard.handle, ctypes.byref(ums), ctypes.byref(unl), ctypes.byref(unb))
            lib.ital_profile_enable(shard.handle, 0)
            upd_ms.append(ums.value / max(1, unl.value))
            upd_gbs.append(unb.value / 1e9 / (ums.value / 1e3) if ums.value > 0 else 0.0)
            fetch_each.append((w1 - w0) * 1e3)
            w3 = w2
            tf += w1 - w0
            tu += w2 - w1
            tt_top += w3 - w2
        learner.top_results(100)                      # first call allocates the sort buffers
        tops = []
        for _ in range(10):
            w0 = time.perf_counter()
            learner.top_results(100)
            tops.append(time.perf_counter() - w0)
        tt_top = float(np.median(tops)) * args.rounds
        rounds = {'rounds': args.rounds, 'fetch_ms': float(np.median(fetch_each)), 'fetch_ms_max': float(np.max(fetch_each)),
                  'update_ms': tu / args.rounds * 1e3,
                  'top_results_100_ms': tt_top / args.rounds * 1e3,
                  'fetch_ms_each': fetch_each,
                  'update_pass_ms': float(np.median(upd_ms)), 'update_pass_GBs': float(np.median(upd_gbs)),
                  'labelled_after': n_lab + args.rounds * args.batch,
                  'note': 'wall clock (fetch: median; the maximum contains the one-off growth of the projection matrix when the '
                          'column capacity doubles); update(%d labels) = one multi-column streaming pass + host bookkeeping; '
                          'top_results(100) = device radix sort of the local means, 100 indices read back' % args.batch}

    used_peer = bool(getattr(learner, '_peer', False))
    learner.close()
    if rank != 0:
        dist.destroy_process_group()
        return
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dev / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'SYN-1M: %d rows per GPU x %d GPU(s), d=%d float32-representable features in HBM, '
                               'fetch_unlabelled(%d), |L|=%d, length_scale=1, perfect user; float64 arithmetic'
                               % (args.rows, world, args.dim, args.batch, n_lab),
                   'rows_per_gpu': args.rows, 'd': args.dim, 'batch': args.batch, 'labelled': n_lab,
                   'candidates_ranked_per_step': ranked, 'l2': 'inputs (2 GB per GPU) exceed the 126 MB L2; no flush',
                   'parallelism': 'rows sharded over %d GPU(s); per greedy step every shard %s' % (
                       world, 'stores its proposal into its peers\' memory over NVLink (CUDA IPC; no NCCL in the loop)'
                       if used_peer else 'contributes one record to an NCCL all-gather'),
                   'batch_selected': [int(i) for i in ret]},
        'e2e': {'value': ranked * args.steps / wall, 'unit': UNIT, 'ms_per_step': wall / args.steps * 1e3,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'note': 'wall clock of ITAL.fetch_unlabelled through the Python/ctypes/C-ABI boundary; the pool '
                        'itself is resident (uploaded once by fit: %.0f ms, %.2f GB)' % (t_fit * 1e3, X.nbytes / 1e9)},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'kernel': 'k_extend_bulk / k_extend (streaming pass: row . z in f64, RBF, projection)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'traffic_source': traffic_src,
                     'peak_source': peak_src, 'launches': int(nl.value),
                     'avg_launch_ms': ms.value / max(1, nl.value),
                     'algorithmic_bytes_per_launch': nbytes.value / max(1, nl.value),
                     'share_of_step': ms.value / 1e3 / dev},
        'exhaustive': exh,
        'lazy_rows': lazy,
        'al_rounds': rounds,
        'fetch_stats_per_step': stats,
        'setup': {'generate_s': t_gen, 'fit_s': t_fit, 'update_9_labels_s': t_update},
    }
    if world == 1 and not args.no_cpu_baseline:
        rows = args.cpu_rows if args.cpu_rows else 12000
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            rate, per, cret, done = oracle_fetch_rate(rows, args.dim, args.batch, 1, steps=1)
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                                'sample': 'first %d rows of the same pool, one fetch_unlabelled(%d) with every '
                                          'candidate scored (%.1f s); numpy float64 oracle, single process'
                                          % (rows, args.batch, per)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
