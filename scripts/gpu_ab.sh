#!/bin/bash
# A/B of search knobs on the C3 bench
mkdir -p gpurun_out
for CFG in "1 0" "0 0" "1 6" "0 6" "1 2"; do
  set -- $CFG
  export MP2P_TILE_STRIDE=$1 MP2P_DESCENT=$2
  timeout 900 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_s$1_d$2.json 2> gpurun_out/bench_c3_s$1_d$2.err
  python - "gpurun_out/bench_c3_s$1_d$2.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'warm', round(d['config'].get('ms_per_step_l2_warm_informative',0),4), 'nn_ms', round(r['kernel_ms'],4), 'probes', r['probes'], 'cands', r['candidates'])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
done
