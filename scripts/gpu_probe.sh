#!/bin/bash
mkdir -p gpurun_out
for W in C2 C3; do timeout 600 python scripts/probe_e2e.py $W 2>&1 | tee gpurun_out/probe_e2e_$W.txt | tail -16; done
