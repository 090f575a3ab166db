#!/bin/bash
# Round-2 GPU call 1: new-path tests first (fail fast), then the whole GPU suite, then the default bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or swarm_randoms or graph_replay or foreign or integration_md or scaled_operand" 2>&1 | tail -25
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^posterior N=" | tail -15
timeout 900 python bench.py > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/r2a_bench_1gpu.err
python - <<'PY'
import json
try:
    j=json.loads(open("gpurun_out/r2a_bench_1gpu.json").read().strip().splitlines()[-1])
    print("C4 step %.3f ms  K2 %.3f ms  frac %.4f  e2e %.3f ms  launches %d parity %s sharded %s" % (j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["parity"], j["sharded_parity"]))
    s=j["secondary"][0]
    print("C5 iter %.3f ms  K2 %.3f ms frac %.3f e2e %.1f ms graph %s parity %s" % (s["ms_per_step"], s["roofline"]["kernel_ms_per_launch"], s["roofline"]["frac"], s["e2e"]["ms_per_step"], s["graph"], s["parity"]))
except Exception as e:
    print("bench parse failed", e)
PY
for c in C2 C3; do timeout 300 python bench.py --config $c --steps 50 --no-sharded-parity > gpurun_out/r2a_bench_$c.json 2> gpurun_out/r2a_bench_$c.err; python - $c <<'PY'
import json,sys
try:
    j=json.loads(open("gpurun_out/r2a_bench_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("%s step %.4f ms  K2 %.4f ms  frac %.4f  e2e %.4f ms  launches %d parity %s" % (sys.argv[1], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["parity"]))
except Exception as e:
    print("bench failed", e)
PY
done
