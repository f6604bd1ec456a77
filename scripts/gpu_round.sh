#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for V in 1 0; do
  MP2P_KNN_LANE=$V timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_lane$V.json 2> gpurun_out/bench_c3_lane$V.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_c3_lane$V.json').read()); r=d['roofline']
print('lane=$V', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:r[k] for k in ('probes','candidates','climbed_queries','kernel_ms')}, r['other_kernels_ms'])"; tail -2 gpurun_out/bench_c3_lane$V.err
done
cp gpurun_out/bench_c3_lane1.json gpurun_out/bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_match_knn_lane" -s 3 -c 1 -f -o gpurun_out/prof_c3 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.json
