#!/bin/bash
# compute-sanitizer over the smoke run (every kernel family once, small sizes): memcheck, racecheck, synccheck
mkdir -p gpurun_out
for T in memcheck racecheck; do
  timeout 150 compute-sanitizer --tool $T --print-limit 20 python scripts/sanitize_run.py > gpurun_out/sanitizer_$T.log 2>&1
  echo "$T rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run OK|Error|hazard" gpurun_out/sanitizer_$T.log | head -8
done
