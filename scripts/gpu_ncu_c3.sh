#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt|k_plane_fit" -s 8 -c 2 -f -o gpurun_out/prof_c3_search4 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_C3s.log 2>&1; echo "full C3 search rc=$?"
