#!/bin/bash
# launch list of the 256-query filter batch after the merge_lists change (ratio merge_lists / scan_tc2 inside one run)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_filter_final.csv python bench.py --steps 8 --warmup 3 --cpu-seconds 0 --sections vector --vector-kernel filt256p > gpurun_out/c24_launch_bench.log 2>&1; echo "launch list rc=$?"
python profiles/summarize_ncu.py list gpurun_out/r02_launches_filter_final.csv 2>/dev/null | grep -E "scan_tc|merge_lists|kth_from|refine|fallback|prep_split|launches" | cut -c1-140
