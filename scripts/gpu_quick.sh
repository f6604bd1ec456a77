#!/bin/bash
# quick visit: GPU tests + C3 and C2 bench lines (resident cloud)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'warm', round(d['config'].get('ms_per_step_l2_warm_informative',0),4), 'e2e', round(d['e2e']['ms_per_step'],4), 'nn_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'probes', r['probes'], 'cands', r['candidates'], r.get('per_query_max'), r['other_kernels_ms'])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
}
timeout 900 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_cloud.json 2> gpurun_out/bench_c3_cloud.err; show gpurun_out/bench_c3_cloud.json; tail -2 gpurun_out/bench_c3_cloud.err
timeout 600 python bench.py --workload C2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_cloud.json 2> gpurun_out/bench_c2_cloud.err; show gpurun_out/bench_c2_cloud.json; tail -2 gpurun_out/bench_c2_cloud.err
