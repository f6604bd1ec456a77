#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1200 gpurun_out/bench_c2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 900 gpurun_out/bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt" -s 3 -c 1 -f -o gpurun_out/prof_c3 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out | tail -5
