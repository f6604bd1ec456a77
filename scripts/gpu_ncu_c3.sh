#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_C3.csv python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_C3.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt|k_plane_fit|k_gn_accumulate|k_compact_pt2pl" -s 12 -c 4 -f -o gpurun_out/prof_c3_all python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_C3s.log 2>&1; echo "full C3 rc=$?"
