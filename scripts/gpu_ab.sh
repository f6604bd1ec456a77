#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for V in "" _mb3; do
  export MP2P_B200_LIB=$PWD/mp2p_icp_b200/libmp2p_b200$V.so
  timeout 900 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_v$V.json 2> gpurun_out/bench_c3_v$V.err
  python - "gpurun_out/bench_c3_v$V.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'warm', round(d['config'].get('ms_per_step_l2_warm_informative',0),4), 'nn_ms', round(r['kernel_ms'],4), r.get('per_query_max'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
done
