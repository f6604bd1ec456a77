#!/bin/bash
# N-GPU visit (gpurun --gpus N): whole GPU suite (includes the torchrun parity worker with both
# transports) + sharded C2/C3 bench with the peer transport
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_n$N.log
bash scripts/gpu_multi2.sh $N C2 C3
