#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pt2ln or gn_ or fused or pt2pl" > gpurun_out/pytest_pt2ln.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_pt2ln.log
bash scripts/gpu_sanitize.sh
