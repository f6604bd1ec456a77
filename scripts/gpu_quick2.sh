#!/bin/bash
# quick visit: GPU tests + C2 and C3 bench lines (no CPU baseline)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for W in C2 C3; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?"; tail -3 gpurun_out/bench_$W.err
  python - $W <<'P'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1]); r=d['roofline']; e=d['e2e']
    print(sys.argv[1],'dev ms',round(d['ms_per_step'],4),'e2e ms',round(e['ms_per_step'],4),'upload',round(e['ms_per_step_pairs_uploaded_again'],4),'fused_host',e['ms_per_step_fused_call_host_cloud'],'nn_ms',round(r['kernel_ms'],4),'frac',round(r['frac'],3),r['other_kernels_ms'])
except Exception as ex: print('unreadable',ex)
P
done
