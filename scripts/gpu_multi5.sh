#!/bin/bash
# N-GPU confirmation (gpurun --gpus N): torchrun parity worker (both transports) + C2 bench, peer transport
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -k "multi_gpu" -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi_n$N.log
for W in C2 C3; do
MP2P_B200_TRANSPORT=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $W > gpurun_out/bench_${W}_n${N}_peer.json 2> gpurun_out/bench_${W}_n${N}_peer.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_${W}_n${N}_peer.json').read().strip().splitlines()[-1]); print('$W N=$N peer value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], 'pairs', d['config']['pairs'], d['config']['collectives'])
except Exception as e:
    print('unreadable', e)
P
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_${W}_n${N}_peer.err | tail -5
done
