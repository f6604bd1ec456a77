#!/bin/bash
# round 2, visit 2: CTA-size A/B of both k > 1 searches on C3, then ncu --set full of the two kernels
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms', round(d['ms_per_step'],4), 'nn_ms', round(r['kernel_ms'],4), 'cands', r['candidates'], 'probes', r['probes'], 'pairs', d['config']['pairs'])
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
P
}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; show gpurun_out/bench_c3_$tag.json; tail -2 gpurun_out/bench_c3_$tag.err
}
run v1_nt128 MP2P_KNN_V1=1 MP2P_KNN_NT=128
run v1_nt64 MP2P_KNN_V1=1 MP2P_KNN_NT=64
run p3_nt128 MP2P_KNN_NT=128
run p3_nt64 MP2P_KNN_NT=64
MP2P_KNN_NT=64 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "knn or pt2pl or pt2ln or adaptive or c2_small" > gpurun_out/pytest_knn_nt64.log 2>&1; echo "nt64 pytest rc=$?"; tail -1 gpurun_out/pytest_knn_nt64.log
for V in 0 1; do
  MP2P_KNN_V1=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_match_pt2pt" -s 4 -c 1 -f -o gpurun_out/prof_r2_knn_v$V python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_knn_v$V.log 2>&1; echo "ncu V1=$V rc=$?"
done
ls -la gpurun_out/*.ncu-rep
