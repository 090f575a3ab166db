#!/bin/bash
# Final single-GPU evidence: full GPU suite, smoke, default bench (C4 + C5 secondary), C2 / C3 lines, reference arm, ncu launch list
# of the default bench command and one full capture of the dominant kernel (traffic).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2q_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2q_smoke.log
timeout 900 python bench.py > gpurun_out/r2q_bench_1gpu.json 2> gpurun_out/r2q_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2q_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2q_bench_reference_arm.json 2> /dev/null
for c in C2 C3; do timeout 300 python bench.py --config $c --steps 50 --no-secondary > gpurun_out/r2q_bench_1gpu_$c.json 2> /dev/null; done
timeout 300 python bench.py --config C3 --fp64 --steps 50 --no-secondary > gpurun_out/r2q_bench_1gpu_C3_fp64.json 2> /dev/null
timeout 600 python bench.py --config C5 --steps 20 > gpurun_out/r2q_bench_1gpu_C5.json 2> /dev/null
python - <<'PY'
import json
for f in ["r2q_bench_1gpu","r2q_bench_1gpu_C2","r2q_bench_1gpu_C3","r2q_bench_1gpu_C3_fp64","r2q_bench_1gpu_C5","r2q_bench_reference_arm"]:
    try:
        j=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=j.get("roofline") or {}
        print("%-28s %s value %.3e step %.4f ms K2 %s frac %s e2e %s parity %s" % (f, j.get("dtype"), j["value"], j["ms_per_step"], r.get("kernel_ms_per_launch"), r.get("frac"), (j.get("e2e") or {}).get("ms_per_step"), {k:(v.get("ok", v.get("equal")) if isinstance(v,dict) else v) for k,v in (j.get("parity") or {}).items()}))
    except Exception as e:
        print(f, "failed", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_posterior_tma -s 2 -c 1 -o gpurun_out/r2q_k2_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-sharded-parity > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_sets_fused -s 3 -c 1 -o gpurun_out/r2q_sets_fused_C4 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-sharded-parity > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_fit_cluster -s 50 -c 1 -o gpurun_out/r2q_fit_cluster_N256 -f python tools/bench_fit.py > /dev/null 2>&1
ls -la gpurun_out/r2q*.ncu-rep
