#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for F in 1 0; do
  export MP2P_FUSED_ITERATION=$F
  timeout 600 python bench.py --workload C2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_f$F.json 2> gpurun_out/bench_c2_f$F.err
  python - "gpurun_out/bench_c2_f$F.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'warm', round(d['config'].get('ms_per_step_l2_warm_informative',0),4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d['gpu_launches'], 'pairs', d['config']['pairs'])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
  tail -2 gpurun_out/bench_c2_f$F.err
done
