#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "posterior_matches or swarm" 2>&1 | tail -3
run() { # name lib rows48
  SAFEOPT_B200_LIB=$2 SO_K2_ROWS48=$3 timeout 300 python bench.py --config C5 --steps 30 --no-cpu-baseline --no-sharded-parity $4 2>gpurun_out/err_$1.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 rows48=$3 $4 step %.4f K2 %.4f frac %.3f'%(j['ms_per_step'], j['roofline']['kernel_ms_per_launch'], j['roofline']['frac']))"
}
for v in 1 0; do
  run main "" $v; run unroll2 $PWD/tools/ab/unroll2.so $v; run nogen $PWD/tools/ab/nogen.so $v
done 2>&1 | tee gpurun_out/rows48b.txt
