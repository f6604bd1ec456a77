#!/bin/bash
# quick N-GPU visit: C2 (and optionally C3) bench with the peer transport only
N=${1:-2}; shift
mkdir -p gpurun_out
for W in "$@"; do
MP2P_B200_TRANSPORT=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $W > gpurun_out/bench_${W}_n${N}_peer.json 2> gpurun_out/bench_${W}_n${N}_peer.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_${W}_n${N}_peer.json').read().strip().splitlines()[-1]); print('$W N=$N peer value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'warm', round(d['config']['ms_per_step_l2_warm_informative'],4), 'launches', d['gpu_launches'], 'pairs', d['config']['pairs'], d['config']['collectives'])
except Exception as e:
    print('unreadable', e)
P
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_${W}_n${N}_peer.err | tail -5
done
