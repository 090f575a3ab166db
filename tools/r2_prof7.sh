#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python tools/bench_fit.py 2>&1 | tail -3
for c in C2 C3; do timeout 300 python bench.py --config $c --steps 30 --no-sharded-parity --no-secondary > gpurun_out/r2k_bench_$c.json 2> gpurun_out/r2k_bench_$c.err; tail -2 gpurun_out/r2k_bench_$c.err; python - $c <<'PY'
import json,sys
try:
    j=json.loads(open("gpurun_out/r2k_bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("%s %s step %.4f ms  K2 %.4f ms x%.0f frac %.4f  e2e %.4f ms  launches %d parity ok=%s" % (sys.argv[1], j["dtype"], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["kernel_launches_per_step"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["parity"]["vs_port"]["ok"]))
except Exception as e:
    print("bench failed", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2k_launches_C3.csv python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-sharded-parity > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2k_launches_C3.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:60], r[7], r[8], r[-1])
PY
