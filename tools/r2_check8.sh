#!/bin/bash
# Round-2 GPU call (N GPUs): the strong-scaled default bench only.
N=${1:-8}
TAG=${2:-r2c}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"; grep -v "^W\|^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_bench_${N}gpu.err | tail -12
python - $N $TAG <<'PY'
import json,sys
try:
    j=json.loads([l for l in open("gpurun_out/%s_bench_%sgpu.json"%(sys.argv[2],sys.argv[1])).read().strip().splitlines() if l.startswith("{")][-1])
    print("C4 x%s step %.3f ms  K2 %.3f ms  frac %.4f  e2e %.3f ms  launches %d exchange=%s\n parity %s\n sharded %s" % (sys.argv[1], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["exchange"], j["parity"], j["sharded_parity"]))
    s=j["secondary"][0]
    print("C5 iter %.3f ms  K2 %.3f ms frac %.3f e2e %.1f ms graph %s exchange=%s" % (s["ms_per_step"], s["roofline"]["kernel_ms_per_launch"], s["roofline"]["frac"], s["e2e"]["ms_per_step"], s["graph"], s["exchange"]))
except Exception as e:
    print("bench parse failed", e)
PY
