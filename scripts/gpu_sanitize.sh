#!/bin/bash
# compute-sanitizer over one small call of every kernel family (scripts/sanitize_run.py): memcheck with the shipped
# defaults, memcheck with the A/B variants switched on (thread-per-query search with most queries handed over,
# tight boxes), racecheck
mkdir -p gpurun_out
run() { tag=$1; tool=$2; shift 2
  env "$@" timeout 400 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py > gpurun_out/sanitizer_$tag.log 2>&1
  echo "$tag rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run OK|Error|hazard|Traceback" gpurun_out/sanitizer_$tag.log | head -8
}
run memcheck memcheck MP2P_UNUSED=1
run memcheck_variants memcheck MP2P_KNN_THREAD=1 MP2P_KNN_DEFER_PROBES=4 MP2P_KNN_DEFER_CANDS=30 MP2P_INDEX_BOX=1
run racecheck racecheck MP2P_UNUSED=1
