#!/bin/bash
# Final single-GPU evidence for the round: full GPU suite, smoke, default bench (C4 + C5 secondary), C2 / C3 / C5 lines, fit times,
# K2 across N, ncu launch list of the default bench command, full captures of the grid kernel at N = 128 / 288 / 512 and of the fit.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -4 | tee gpurun_out/r2u_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2u_smoke.log
timeout 900 python bench.py > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2u_bench_1gpu.err
for c in C2 C3; do timeout 300 python bench.py --config $c --steps 50 --no-secondary > gpurun_out/r2u_bench_1gpu_$c.json 2> /dev/null; done
timeout 300 python bench.py --config C3 --fp64 --steps 50 --no-secondary > gpurun_out/r2u_bench_1gpu_C3_fp64.json 2> /dev/null
timeout 600 python bench.py --config C5 --steps 20 > gpurun_out/r2u_bench_1gpu_C5.json 2> /dev/null
python - <<'PY'
import json
for f in ["r2u_bench_1gpu","r2u_bench_1gpu_C2","r2u_bench_1gpu_C3","r2u_bench_1gpu_C3_fp64","r2u_bench_1gpu_C5"]:
    try:
        j=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=j.get("roofline") or {}
        print("%-24s %s value %.3e step %.4f ms K2 %s frac %s e2e %s parity %s" % (f, j.get("dtype"), j["value"], j["ms_per_step"], r.get("kernel_ms_per_launch"), r.get("frac"), (j.get("e2e") or {}).get("ms_per_step"), {k:(v.get("ok", v.get("equal")) if isinstance(v,dict) else v) for k,v in (j.get("parity") or {}).items()}))
        for s in j.get("secondary", []):
            print("   secondary %s: iter %.4f ms  K2 %.4f frac %.3f e2e %.1f ms parity %s" % (s["name"], s["ms_per_step"], s["roofline"]["kernel_ms_per_launch"], s["roofline"]["frac"], s["e2e"]["ms_per_step"], s["parity"]))
    except Exception as e:
        print(f, "failed", e)
PY
python tools/bench_fit.py 2>/dev/null | tail -1 | tee gpurun_out/r2u_fit_times.json
python tools/time_k2_vs_n.py 32 64 96 128 192 256 257 280 281 288 320 384 416 512 640 2>/dev/null | grep "^{" | tee gpurun_out/r2u_k2_vs_n.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for n in 128 288 512; do
  ncu --set full --import-source on --clock-control none -k regex:k_posterior -s 3 -c 1 -o gpurun_out/r2u_k2_N$n -f python tools/time_k2_vs_n.py $n > /dev/null 2>&1
done
ncu --set full --import-source on --clock-control none -k regex:k_fit_cluster -s 50 -c 1 -o gpurun_out/r2u_fit_cluster_N256 -f python tools/bench_fit.py > /dev/null 2>&1
ls -la gpurun_out/r2u*.ncu-rep
