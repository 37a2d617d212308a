#!/bin/bash
# Compare TMA-staged streaming-kernel variants (ital_b200/lib/variant_bulk_*.so) with the default build.
for lib in ital_b200/lib/libital_b200.so ital_b200/lib/variant_bulk_*.so; do
  ITAL_B200_BULK=1 ITAL_B200_LIB=$PWD/$lib timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --exhaustive-steps 0 --lazy-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib', 'fetch_ms=%.4f'%d['ms_per_step'], 'extend_ms=%.4f'%r['avg_launch_ms'], 'GB/s=%.0f'%r['achieved'], 'frac=%.3f'%r['frac'])"
done
