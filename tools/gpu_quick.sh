#!/bin/bash
# Quick GPU visit: parity tests, short bench, optionally one full ncu capture of the trace kernel (per-line export happens on the CPU side).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_quick.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], [(a["precision"], round(a["value"], 1)) for a in d.get("alt_modes", [])])
PY
if [ "$1" = "ncu" ]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt"
  ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/prof_trace $CMD > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
