#!/bin/bash
# Compare k_extend tuning variants (built as ital_b200/lib/variant_*.so) on the bench workload.
for lib in ital_b200/lib/libital_b200.so ital_b200/lib/variant_*.so; do
  ITAL_B200_LIB=$PWD/$lib timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --exhaustive-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib', 'fetch_ms=%.4f'%d['ms_per_step'], 'extend_ms=%.4f'%r['avg_launch_ms'], 'GB/s=%.0f'%r['achieved'], 'frac=%.3f'%r['frac'])"
done
