#!/bin/bash
# One GPU-box visit: parity tests, bench (own arm + reference arm), then -- with "ncu" -- the launch list and one full
# capture each of the default trace kernel (plain layout), the regroup layout and the all-fp64 kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
if [ "$1" = "ncu" ]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  for v in "prof_trace:" "prof_trace_regroup:--layout regroup" "prof_trace_brute:--precision fp64-brute"; do
    name=${v%%:*}; extra=${v#*:}
    $CMD $extra > gpurun_out/plain_$name.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/$name $CMD $extra > gpurun_out/ncu_$name.log 2>&1
    tail -2 gpurun_out/ncu_$name.log
  done
fi
