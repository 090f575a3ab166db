#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "golden or posterior or grid or shared or c3 or fused or bo_loop" 2>&1 | tail -4
for v in 1 0; do echo "SO_K2_SMALLN=$v"; SO_K2_SMALLN=$v python tools/time_k2_vs_n.py 32 64 96 128 2>&1 | grep "^{"; done
for v in 1 0; do SO_K2_SMALLN=$v timeout 300 python bench.py --config C3 --fp64 --steps 50 --no-secondary --no-cpu-baseline --no-sharded-parity 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 fp64 smalln=$v step %.4f K2 %.4f frac %.3f'%(j['ms_per_step'], j['roofline']['kernel_ms_per_launch'], j['roofline']['frac']))"; done
for v in 1 0; do SO_K2_SMALLN=$v timeout 300 python bench.py --config C2 --steps 50 --no-secondary --no-cpu-baseline --no-sharded-parity 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 smalln=$v step %.4f K2 %.4f frac %.3f'%(j['ms_per_step'], j['roofline']['kernel_ms_per_launch'], j['roofline']['frac']))"; done
