#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi.sh N [what...]: what = parity (the torchrun parity worker, owner-partitioned
# claims on and off), bench (the default line at N GPUs: C3 weak scaling + the C5 strong-scaling object), c5 (C5 alone)
N=$1; shift
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d.get("kernel_ms_rank0"), d["config"]["workload"], d["scaling"], 'N', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'collectives', d['config']['collectives'], 'parity', d.get('parity_vs_n1'))
    c=d.get('c5') or {}
    if c: print('  c5:', {k:c.get(k) for k in ('n_gpus','ms_per_step','value','queries_per_gpu','gpu_launches_per_step','parity_vs_n1','error')})
except Exception as e:
    print('unreadable', e)
P
}
for what in "$@"; do
  case $what in
    parity)
      for oc in 1 0; do
        MP2P_B200_OWNER_CLAIMS=$oc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/multi_gpu_parity_worker.py > gpurun_out/parity_n${N}_oc$oc.log 2>&1; echo "parity N=$N owner_claims=$oc rc=$?"; grep -E "world=|pt2pl" gpurun_out/parity_n${N}_oc$oc.log; tail -3 gpurun_out/parity_n${N}_oc$oc.log | grep -iE "error|Traceback" 
      done ;;
    bench)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; show gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err | grep -v OMP ;;
    c5)
      for oc in 1 0; do
        MP2P_B200_OWNER_CLAIMS=$oc timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload C5 > gpurun_out/bench_c5_n${N}_oc$oc.json 2> gpurun_out/bench_c5_n${N}_oc$oc.err; echo "C5 N=$N owner_claims=$oc rc=$?"; show gpurun_out/bench_c5_n${N}_oc$oc.json; tail -3 gpurun_out/bench_c5_n${N}_oc$oc.err | grep -v OMP
      done ;;
    c5n1)
      timeout 900 python bench.py --steps 10 --warmup 3 --workload C5 --no-cpu-baseline > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "C5 N=1 rc=$?"; show gpurun_out/bench_c5_n1.json ;;
  esac
done
