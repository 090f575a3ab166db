#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp32 or fit or fused" 2>&1 | tail -5
python tools/bench_fit.py 2>&1 | tail -2
for mc in 1 0; do SO_F32_MULTICAST=$mc timeout 300 python bench.py --config C3 --steps 30 --no-sharded-parity --no-secondary --no-cpu-baseline > gpurun_out/r2o_bench_C3_mc$mc.json 2> gpurun_out/r2o_bench_C3_mc$mc.err; tail -2 gpurun_out/r2o_bench_C3_mc$mc.err; python - $mc <<'PY'
import json,sys
try:
    j=json.loads(open("gpurun_out/r2o_bench_C3_mc%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("C3 multicast=%s %s step %.4f ms  K2 %.4f ms frac %.4f  e2e %.4f ms" % (sys.argv[1], j["dtype"], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"]))
except Exception as e:
    print("bench failed", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2o_launches_C3.csv python bench.py --config C3 --steps 3 --warmup 1 --no-cpu-baseline --no-sharded-parity > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2o_launches_C3.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]:
    print(r[4][:60], r[7], r[8], r[-1])
PY
timeout 300 python bench.py --config C2 --steps 30 --no-sharded-parity --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 step %.4f K2 %.4f e2e %.4f'%(j['ms_per_step'], j['roofline']['kernel_ms_per_launch'], j['e2e']['ms_per_step']))"
