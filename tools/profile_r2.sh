#!/bin/bash
# ncu captures behind profiles/r02_*: run on the GPU box (gpurun -- bash tools/profile_r2.sh).  The .ncu-rep files are
# condensed on the box (tools/summarize_ncu.py) and removed: gpurun_out/ carries at most 64 MiB back.  Numbers printed
# by runs under ncu are never bench values.
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_r2a.csv python bench.py --steps 3 --warmup 3 \
    --no-cpu-baseline --secondary-steps 2 --exhaustive-steps 0 --general-steps 0 --model-steps 0 --rounds 2 --no-strong > $O/launches_r2a.log 2>&1
python tools/summarize_ncu.py launches $O/launches_r2a.csv > $O/r02_launches_fetch.txt
cap() {   # name, kernel regex, skip, count, driver mode
    $NCU --set full --import-source on -k "regex:$2" -s $3 -c $4 -o $O/prof_$1 -f python tools/probe/prof_driver.py $5 > $O/prof_$1.log 2>&1
    python tools/summarize_ncu.py full $O/prof_$1.ncu-rep > $O/r02_$1_full.txt
    ncu -i $O/prof_$1.ncu-rep --page details --csv 2>/dev/null | grep -E "Stall|stall|Pipe|Issue|Eligible|Theoretical Occupancy|Achieved Occupancy" | head -60 > $O/r02_$1_details.csv
    rm -f $O/prof_$1.ncu-rep
}
cap k_fetch_fused k_fetch_fused 2 2 fused
cap k_eval "^k_eval$" 0 3 exhaustive
cap k_eval_general k_eval_general 0 3 general
cap k_extend_bulk_multi k_extend_bulk_multi 0 2 update
cap k_extend_bulk "^k_extend_bulk$" 1 2 streaming
ls -la $O | tail -20
