#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do timeout 1500 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -3; done
timeout 300 python bench.py --config C2 --steps 50 --no-sharded-parity --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 step %.4f K2 %.4f e2e %.4f'%(j['ms_per_step'], j['roofline']['kernel_ms_per_launch'], j['e2e']['ms_per_step']))"
timeout 300 python tools/time_step_multi.py 2>&1 | grep "^rank"
