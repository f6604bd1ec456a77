#!/bin/bash
# N-GPU visit (gpurun --gpus N): torchrun parity worker + sharded bench at N and at 1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_parity.py -k "multi_gpu or sharded" -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_multi.log
for W in C2 C3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $W > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/bench_${W}_n$N.json').read().strip().splitlines()[-1]); print('$W N=$N value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], 'pairs', d['config']['pairs'])
except Exception as e:
    print('unreadable', e)
P
tail -3 gpurun_out/bench_${W}_n$N.err
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_C2_n1.json 2> gpurun_out/bench_C2_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_C2_n1.json').read().strip().splitlines()[-1]); print('C2 N=1 value', round(d['value'],1), 'ms', round(d['ms_per_step'],4))"
