#!/bin/bash
# One GPU call: full GPU test suite, the default bench twice (step / K2 / e2e), the config-5 swarm tool.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for i in 1 2; do
  python bench.py --no-cpu-baseline --steps 10 > gpurun_out/chk_bench_$i.json 2> gpurun_out/chk_bench_$i.err || tail -3 gpurun_out/chk_bench_$i.err
  python - $i <<'PY'
import json,sys
try:
    j=json.loads(open("gpurun_out/chk_bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("step %.3f ms  K2 %.3f ms  frac %.4f  e2e %.3f ms  launches %d" % (j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["gpu_launches"]))
except Exception as e:
    print("bench failed", e)
PY
done
python tools/bench_swarm.py --no-cpu > gpurun_out/chk_swarm.json 2> gpurun_out/chk_swarm.err || tail -3 gpurun_out/chk_swarm.err
cut -c1-700 gpurun_out/chk_swarm.json
