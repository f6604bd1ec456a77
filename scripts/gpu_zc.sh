#!/bin/bash
# zero-copy visit: GPU suite, e2e probe with and without zero copy, C2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for Z in 1 0; do
  echo "== MP2P_ZERO_COPY=$Z"
  MP2P_ZERO_COPY=$Z timeout 300 python scripts/probe_e2e.py C2 2>&1 | tee gpurun_out/probe_e2e_C2_zc$Z.txt | head -9
done
timeout 600 python bench.py --workload C2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_C2.json').read().strip().splitlines()[-1]); e=d['e2e']
print('C2 value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],4), 'floor', round(e['pcie_floor_ms'],4))
P
tail -3 gpurun_out/bench_C2.err
