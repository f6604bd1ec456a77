#!/bin/bash
# A/B of the k > 1 search (round 2): MP2P_KNN_V1=1 (per-run scan of round 1) vs the dense-list search with
# pruned descent, 2 or 3 phases per level; parity subset + C3 bench (no extras) per setting.
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'parts', {k:round(v,4) for k,v in r['kernel_ms_parts'].items() if v}, 'cands/q', round(r['candidates_per_query'],1), 'probes/q', round(r['probes_per_query'],1), 'max', r['per_query_max'], 'pairs', d['config']['pairs'], 'frac', round(r['frac'],3))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or pt2pl or pt2ln or adaptive or c2_small" > gpurun_out/pytest_knn_$tag.log 2>&1; echo "$tag pytest rc=$?"; tail -1 gpurun_out/pytest_knn_$tag.log
  env "$@" timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; show gpurun_out/bench_c3_$tag.json; tail -2 gpurun_out/bench_c3_$tag.err
}
for tag in "$@"; do
  case $tag in
    thr) run thr MP2P_KNN_THREAD=1 ;;
    grp) run grp MP2P_KNN_THREAD=0 ;;
    ncut) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_match_knn_thread" -s 4 -c 1 -f -o gpurun_out/prof_r2_knn_thread python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_r2_knn_thread.log 2>&1; echo "ncut rc=$?" ;;
    p3) run p3 MP2P_KNN_PHASES=3 ;;
    p2) run p2 MP2P_KNN_PHASES=2 ;;
    v1) run v1 MP2P_KNN_V1=1 ;;
    nolpt) run nolpt MP2P_KNN_LPT=0 ;;
    v1nolpt) run v1nolpt MP2P_KNN_V1=1 MP2P_KNN_LPT=0 ;;
    fast:*) tag2=${tag#fast:}; IFS=, read -ra envs <<< "$tag2"; env "${envs[@]}" timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c3_fast.json 2> gpurun_out/bench_c3_fast.err; echo "fast $tag2"; show gpurun_out/bench_c3_fast.json; tail -2 gpurun_out/bench_c3_fast.err ;;
    ncu) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt" -s 4 -c 1 -f -o gpurun_out/prof_r2_knn_v3 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_r2_knn_v3.log 2>&1; echo "ncu rc=$?" ;;
    trace) MP2P_KNN_TRACE=1 timeout 300 python scripts/knn_trace.py > gpurun_out/knn_trace_v4.txt 2>&1; tail -8 gpurun_out/knn_trace_v4.txt ;;
    filter) timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter_decimate or covariance" > gpurun_out/pytest_filter.log 2>&1; echo "filter pytest rc=$?"; tail -5 gpurun_out/pytest_filter.log ;;
    full) timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log ;;
  esac
done
