#!/bin/bash
# Round-2 GPU call 2 (N GPUs): the NCCL / peer-exchange tests and the strong-scaled bench at N ranks.
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_gpu" 2>&1 | tail -15 | tee gpurun_out/r2b_two_gpu_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/r2b_bench_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err; echo "bench rc=$?"; grep -v "^W\|^\[W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r2b_bench_${N}gpu.err | tail -12
python - $N <<'PY'
import json,sys
try:
    j=json.loads([l for l in open("gpurun_out/r2b_bench_%sgpu.json"%sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("C4 x%s step %.3f ms  K2 %.3f ms  frac %.4f  e2e %.3f ms  launches %d exchange=%s\n parity %s\n sharded %s" % (sys.argv[1], j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["exchange"], j["parity"], j["sharded_parity"]))
    s=j["secondary"][0]
    print("C5 iter %.3f ms  K2 %.3f ms frac %.3f e2e %.1f ms graph %s exchange=%s" % (s["ms_per_step"], s["roofline"]["kernel_ms_per_launch"], s["roofline"]["frac"], s["e2e"]["ms_per_step"], s["graph"], s["exchange"]))
except Exception as e:
    print("bench parse failed", e)
PY
