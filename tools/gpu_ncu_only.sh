#!/bin/bash
# One full ncu capture of the dominant kernel on the bench command (after a plain run of the same command).
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt $@"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/prof_trace $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
