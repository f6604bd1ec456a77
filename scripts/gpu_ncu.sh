#!/bin/bash
# single-GPU profiling visit: launch lists (durations only) + one --set full capture per hot kernel
mkdir -p gpurun_out
for W in C2 C3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$W.log 2>&1
  echo "launch list $W rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt_nn1|k_compact_pt2pt|k_horn_moments" -s 6 -c 3 -f -o gpurun_out/prof_c2 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_C2.log 2>&1; echo "full C2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt|k_plane_fit|k_gn_accumulate" -s 6 -c 3 -f -o gpurun_out/prof_c3 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_C3.log 2>&1; echo "full C3 rc=$?"
ls -la gpurun_out/*.ncu-rep
