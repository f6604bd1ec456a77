#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m pytest tests/test_gpu_parity.py -k multi_gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -c 1500 gpurun_out/bench_c2_n2.json; tail -5 gpurun_out/bench_c2_n2.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; tail -c 400 gpurun_out/bench_c2_n1.json
