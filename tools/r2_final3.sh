#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -12 | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2v_bench_1gpu.json 2> gpurun_out/r2v_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2v_bench_1gpu.err
for c in C2 C3; do timeout 300 python bench.py --config $c --steps 50 --no-secondary > gpurun_out/r2v_bench_1gpu_$c.json 2> /dev/null; done
python - <<'PY'
import json
for f in ["r2v_bench_1gpu","r2v_bench_1gpu_C2","r2v_bench_1gpu_C3"]:
    try:
        j=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=j.get("roofline") or {}
        print("%-22s %s value %.3e step %.4f ms K2 %.4f frac %.3f e2e %.4f launches %d parity %s" % (f, j.get("dtype"), j["value"], j["ms_per_step"], r.get("kernel_ms_per_launch"), r.get("frac"), (j.get("e2e") or {}).get("ms_per_step"), j["gpu_launches"], {k:(v.get("ok", v.get("equal")) if isinstance(v,dict) else v) for k,v in (j.get("parity") or {}).items()}))
        for s in j.get("secondary", []):
            print("   secondary %s: iter %.4f ms  K2 %.4f frac %.3f e2e %.1f ms graph %s" % (s["name"], s["ms_per_step"], s["roofline"]["kernel_ms_per_launch"], s["roofline"]["frac"], s["e2e"]["ms_per_step"], s["graph"]))
    except Exception as e:
        print(f, "failed", e)
PY
