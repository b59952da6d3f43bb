#!/bin/bash
# compute-sanitizer evidence: every device path under memcheck, racecheck, initcheck and synccheck.
# Writes gpurun_out/sanitizer_<tool>.log; tools/summarize_sanitizer.py turns them into profiles/rNN_sanitizer.txt.
mkdir -p gpurun_out
python tools/sanitize_run.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -20 gpurun_out/sanitize_plain.log; echo "sanitize_run failed WITHOUT the sanitizer"; exit 1; }
for tool in memcheck racecheck initcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok" gpurun_out/sanitizer_$tool.log | tail -3
done
