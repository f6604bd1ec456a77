#!/bin/bash
# A/B of the K = 1 search kernel variants (MP2P_NN1_VARIANT) on C2: parity subset + bench per variant
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'warm', round(d['config'].get('ms_per_step_l2_warm_informative',0),4), 'e2e', round(d['e2e']['ms_per_step'],4), 'nn_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'compact', round(r['other_kernels_ms']['compact'],4))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
}
for V in "$@"; do
  export MP2P_NN1_VARIANT=$V
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pt2pt and not multi_gpu" > gpurun_out/pytest_nn1_v$V.log 2>&1; echo "variant $V pytest rc=$?"; tail -1 gpurun_out/pytest_nn1_v$V.log
  timeout 600 python bench.py --workload C2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_v$V.json 2> gpurun_out/bench_c2_v$V.err; show gpurun_out/bench_c2_v$V.json; tail -2 gpurun_out/bench_c2_v$V.err
done
