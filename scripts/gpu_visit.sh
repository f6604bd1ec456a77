#!/bin/bash
# one single-GPU visit of round 2: the GPU suite, smoke(), the default bench line (C3 + align objects + C5), C2, the
# reference arm, ncu launch lists and one --set full capture per dominant kernel (+ traffic json).  Args: what...
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tests) timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log ;;
    bench) timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"; python scripts/show_bench.py gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err ;;
    c2) timeout 600 python bench.py --workload C2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench C2 rc=$?"; python scripts/show_bench.py gpurun_out/bench_c2.json ;;
    ref) timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err; echo "ref C3 rc=$?"; tail -c 400 gpurun_out/bench_ref_c3.json ;;
    ncu)
      for W in C3 C2; do
        timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_$W.log 2>&1; echo "launch list $W rc=$?"
      done
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt|k_plane_fit|k_gn_loop|k_compact_pt2pl" -s 16 -c 4 -f -o gpurun_out/prof_c3 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_C3.log 2>&1; echo "full C3 rc=$?"
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_iterate_nn1_horn|k_match_pt2pt_nn1|k_compact_pt2pt" -s 8 -c 3 -f -o gpurun_out/prof_c2 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_C2.log 2>&1; echo "full C2 rc=$?" ;;
  esac
done
