#!/bin/bash
# N-GPU visit (gpurun --gpus N): torchrun parity worker + sharded bench at N with both transports
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_parity.py -k "multi_gpu" -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest(peer) rc=$?"; tail -15 gpurun_out/pytest_multi.log
for T in peer nccl; do
for W in C2 C3; do
MP2P_B200_TRANSPORT=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $W > gpurun_out/bench_${W}_n${N}_$T.json 2> gpurun_out/bench_${W}_n${N}_$T.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_${W}_n${N}_$T.json').read().strip().splitlines()[-1]); print('$W N=$N $T value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], 'pairs', d['config']['pairs'], d['config']['collectives'])
except Exception as e:
    print('unreadable', e)
P
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/bench_${W}_n${N}_$T.err | tail -5
done
done
