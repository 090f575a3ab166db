#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=long -k "device_rng_runs" > gpurun_out/r2r_dbg1.log 2>&1; tail -60 gpurun_out/r2r_dbg1.log | cut -c1-220
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/r2r_dbg2.log 2>&1; tail -30 gpurun_out/r2r_dbg2.log | cut -c1-220
