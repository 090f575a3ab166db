#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "fp32" 2>&1 | tail -25
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_sets_fused -s 4 -c 1 -o gpurun_out/r2e_sets_fused_C2 -f python bench.py --config C2 --steps 3 --warmup 1 --no-cpu-baseline --no-sharded-parity > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
