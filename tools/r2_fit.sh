#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "fit or golden or bo_loop" 2>&1 | tail -3
SAFEOPT_B200_LIB=$PWD/tools/ab/stamps.so python tools/bench_fit.py 2> gpurun_out/fit_stamps_err.txt | tail -1
for n in 64 128 256 512; do grep "fit stamps N=$n\]" gpurun_out/fit_stamps_err.txt | tail -1; done > gpurun_out/fit_stamps.txt
cat gpurun_out/fit_stamps.txt | cut -c1-900
echo main; python tools/bench_fit.py 2>/dev/null | tail -1
