#!/bin/bash
# A/B of the pageable-buffer path (hostcopy.hpp): helper threads beside the caller, C3 and C2
mkdir -p gpurun_out
for W in ${WORKLOADS:-C3 C2}; do
for T in ${HELPERS:-0 1 2 3 7}; do
  MP2P_HOST_COPY_THREADS=$T timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_page_${W}_$T.json 2> gpurun_out/bench_page_${W}_$T.err
  python - gpurun_out/bench_page_${W}_$T.json $W $T <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
    print(sys.argv[2], 'helpers', sys.argv[3], 'pinned e2e ms', round(e['ms_per_step'],4), 'pageable ms', round(e['pageable']['ms_per_step'],4))
except Exception as ex:
    print('unreadable', ex)
P
done
done
nproc; lscpu | grep -i "model name\|^CPU(s)\|Thread\|Socket" | head -5
