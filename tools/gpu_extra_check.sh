#!/bin/bash
# Secondary configurations and hazards: bench lines for configs 2 and 3 and the explicit-rows path of config 4, racecheck on the
# kernels that synchronise through shared memory without TMA.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in C2 C3; do
  python bench.py --config $c --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/x_bench_$c.json 2> gpurun_out/x_bench_$c.err || tail -3 gpurun_out/x_bench_$c.err
done
python bench.py --explicit-rows --no-cpu-baseline --steps 10 > gpurun_out/x_bench_C4_rows.json 2> gpurun_out/x_bench_C4_rows.err || tail -3 gpurun_out/x_bench_C4_rows.err
python - <<'PY'
import json
for n in ["C2","C3","C4_rows"]:
    try:
        j=json.loads(open("gpurun_out/x_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.4g evals/s  step %.4f ms  K2 %.4f ms x %d GPs  e2e %.4f ms  launches/step %.1f" % (j["value"], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["config"]["n_gps"], j["e2e"]["ms_per_step"], j["gpu_launches"]/j["steps"]))
    except Exception as e:
        print(n, "failed", e)
PY
if [ "$1" = "--racecheck" ]; then
  timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests -m gpu -q -x -k "chained or safeset_kernels or fit_append_and_remove_match_refit[5 or swarm_step or expander_g1 or lipschitz_g1 or fit_matches_lapack[100" > gpurun_out/racecheck.log 2>&1
  grep -n "=========" gpurun_out/racecheck.log | grep -v "Host Frame" | head -20; tail -3 gpurun_out/racecheck.log
fi
