#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp32" 2>&1 | tail -4
timeout 300 python bench.py --config C3 --steps 30 --no-sharded-parity > gpurun_out/r2j_bench_C3.json 2> gpurun_out/r2j_bench_C3.err; tail -2 gpurun_out/r2j_bench_C3.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2j_bench_C3.json").read().strip().splitlines()[-1])
print("C3 %s step %.4f ms  K2 %.4f ms frac %.4f  e2e %.4f ms" % (j["dtype"], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2j_launches_C3.csv python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-sharded-parity > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j_launches_C3.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-16:]:
    print(r[4][:60], r[7], r[8], r[-1])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fit_cluster -c 40 --csv --log-file gpurun_out/r2j_launches_fit.csv python tools/bench_fit.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j_launches_fit.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[::6][:8]:
    print(r[4][:40], r[7], r[8], r[-1])
PY
ncu --set full --import-source on --clock-control none -k regex:k_fit_cluster -s 30 -c 1 -o gpurun_out/r2j_fit_cluster_N256 -f python tools/bench_fit.py > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_posterior_f32 -s 3 -c 1 -o gpurun_out/r2j_f32_C3 -f python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-sharded-parity > /dev/null 2>&1
ls -la gpurun_out/r2j*.ncu-rep
