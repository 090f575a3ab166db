#!/bin/bash
# Two-GPU evidence: the multi-GPU parity tests (skipped on a one-GPU box) and the strong-scaled default bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "two_gpu or 2gpu or multi_gpu or sharded or distributed" 2>&1 | tail -5 | tee gpurun_out/r2v_two_gpu_pytest.log
bash tools/r2_check8.sh 2 r2v
