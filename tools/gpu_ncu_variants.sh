#!/bin/bash
# Tuning aid: instruction count, duration and the main stall reasons of the trace kernel for every library variant in build/variants/.
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
for so in build/variants/*.so; do
  TRAY_LIB=$PWD/$so ncu --metrics $M --clock-control none -k regex:trace_kernel -s 4 -c 1 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
print('$so', {r[-3].replace('smsp__','').replace('average_warps_issue_stalled_','st_')[:34]: r[-1] for r in rows})
"
done
