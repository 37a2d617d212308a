"""Condense ncu output into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/summarize_ncu.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0].replace('void ', '').replace('italk::', '')
        val = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        val = val / 1e3 if unit == 'ns' else (val * 1e3 if unit == 'ms' else val)
        seq.append((name, val))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    total = sum(v[1] for v in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches): %s' % path)
    print('# all launches in the capture: %d kernels, %.1f us' % (len(seq), total))
    print('%-34s %6s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-34s %6d %12.1f %10.1f %7.3f' % (k[:34], v[0], v[1], v[1] / v[0], v[1] / total))
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith('k_score0')]
    if starts:
        # the last fetch of the timed region (the very last one in a bench run is the stepwise fetch that collects
        # the per-step statistics), up to and including the k_mask_all that ends it
        first = starts[-2] if len(starts) >= 2 else starts[-1]
        last = seq[first:]
        for j, (n, _) in enumerate(last):
            if n.startswith('k_mask_all'):
                last = last[:j + 1]
                break
        tot = sum(v for _, v in last)
        print('\n# the last fetch_unlabelled in the capture, launch by launch (%d launches, %.1f us):' % (len(last), tot))
        for n, v in last:
            print('%-34s %10.1f' % (n[:34], v))
        ext = sum(v for n, v in last if n.startswith('k_extend'))
        print('# share of k_extend in that fetch: %.3f' % (ext / tot))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print('# ncu --set full --clock-control none: %s (%d launches captured)' % (path, len(data)))
    name_i = hdr.index('Kernel Name')
    for r in data:
        print('# kernel: %s' % r[name_i][:100])
    for mname in FULL_METRICS:
        if mname in hdr:
            i = hdr.index(mname)
            print('%-88s %-14s %s' % (mname, units[i], '  '.join(r[i] for r in data)))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
