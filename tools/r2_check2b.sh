#!/bin/bash
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/time_step_multi.py 2>&1 | grep "^rank" | tee gpurun_out/r2m_step_split_${N}gpu.txt
timeout 300 python tools/time_step_multi.py 2>&1 | grep "^rank"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp32" 2>&1 | tail -3
