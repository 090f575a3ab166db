#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "golden or posterior or grid or shared or c3 or fused or bo_loop" 2>&1 | tail -4
for v in 1 0; do echo "SO_K2_NS6=$v"; SO_K2_NS6=$v python tools/time_k2_vs_n.py 257 264 281 288 304 320 352 384 2>&1 | grep "^{"; done | tee gpurun_out/ns6.txt
