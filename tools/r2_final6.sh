#!/bin/bash
# Last call of the round: full GPU suite, smoke, default bench on the final code.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -4 | tee gpurun_out/r2x_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2x_smoke.log
timeout 600 python bench.py > gpurun_out/r2x_bench_1gpu.json 2> gpurun_out/r2x_bench_1gpu.err; echo "bench rc=$?"
python tools/bench_fit.py 2>/dev/null | tail -1 | tee gpurun_out/r2x_fit_times.json
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2x_bench_1gpu.json").read().strip().splitlines()[-1])
r=j["roofline"]; s=j["secondary"][0]
print("C4 step %.4f K2 %.4f frac %.4f e2e %.4f parity %s | C5 iter %.4f K2 %.4f frac %.3f e2e %.1f parity %s | cpu %s" % (j["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], j["e2e"]["ms_per_step"], j["parity"]["vs_port"].get("ok"), s["ms_per_step"], s["roofline"]["kernel_ms_per_launch"], s["roofline"]["frac"], s["e2e"]["ms_per_step"], s["parity"]["vs_port"]["ok"], j["cpu_baseline"]["value"]))
PY
