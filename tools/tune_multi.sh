#!/bin/bash
# Compare variants of the multi-column labelled pass (ital_b200/lib/variant_multi_*.so) with the default build.
for lib in ital_b200/lib/libital_b200.so ital_b200/lib/variant_multi_*.so; do
  echo "== $lib"
  ITAL_B200_LIB=$PWD/$lib timeout 300 python tools/probe/multi_time.py 2>&1 | tail -2
done
