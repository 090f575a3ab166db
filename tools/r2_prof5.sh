#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python tools/time_sets.py 2>&1 | tail -20
python tools/bench_fit.py 2>&1 | tail -8
for c in C2 C3 C4; do timeout 300 python bench.py --config $c --steps 30 --no-sharded-parity --no-secondary > gpurun_out/r2i_bench_$c.json 2> gpurun_out/r2i_bench_$c.err; tail -2 gpurun_out/r2i_bench_$c.err; python - $c <<'PY'
import json,sys
try:
    j=json.loads(open("gpurun_out/r2i_bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("%s %s step %.4f ms  K2 %.4f ms x%.0f frac %.4f  e2e %.4f ms  launches %d parity %s" % (sys.argv[1], j["dtype"], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["kernel_launches_per_step"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["parity"]))
except Exception as e:
    print("bench failed", e)
PY
done
