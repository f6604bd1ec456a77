#!/bin/bash
# A/B of library variants on the C3 (and C2) bench: MP2P_B200_LIB selects the .so
mkdir -p gpurun_out
for V in "" _mb5 _mb6 _mb8; do
  export MP2P_B200_LIB=$PWD/mp2p_icp_b200/libmp2p_b200$V.so
  timeout 900 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_v$V.json 2> gpurun_out/bench_c3_v$V.err
  python - "gpurun_out/bench_c3_v$V.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'warm', round(d['config'].get('ms_per_step_l2_warm_informative',0),4), 'nn_ms', round(r['kernel_ms'],4))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
done
