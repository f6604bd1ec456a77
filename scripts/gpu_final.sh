#!/bin/bash
# closing single-GPU visit: whole GPU suite, smoke, short C2 / C3 bench lines (regression check), sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for W in C2 C3; do
timeout 400 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_final.json 2> gpurun_out/bench_${W}_final.err
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_${W}_final.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$W value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'nn_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), r['other_kernels_ms'])
except Exception as e:
    print('unreadable', e)
P
tail -2 gpurun_out/bench_${W}_final.err
done
bash scripts/gpu_sanitize.sh
