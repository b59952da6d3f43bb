#!/bin/bash
# ncu captures only (no tests / bench lines): launch list + full captures of the default, regroup-layout and all-fp64 trace kernels.
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
for v in "prof_trace:" "prof_trace_regroup:--layout regroup" "prof_trace_brute:--precision fp64-brute"; do
  name=${v%%:*}; extra=${v#*:}
  $CMD $extra > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/$name $CMD $extra > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
done
