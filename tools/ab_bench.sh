#!/bin/bash
# A/B of library variants (tools/ab/*.so, built by tools/build_variant.sh) on the default bench: step and K2 times.
# usage: tools/ab_bench.sh [rounds] [variant names...]   (default: every .so under tools/ab)
cd "$(dirname "$0")/.."
rounds=${1:-2}; shift
names=${@:-$(ls tools/ab/*.so | xargs -n1 basename | sed 's/\.so$//')}
for i in $(seq $rounds); do
  for v in $names; do
    SAFEOPT_B200_LIB=$PWD/tools/ab/$v.so python bench.py --no-cpu-baseline --steps 10 > /tmp/ab_$v.json 2> /tmp/ab_$v.err || tail -3 /tmp/ab_$v.err
    python - $v <<'PY'
import json,sys
try:
    j=json.loads(open("/tmp/ab_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "step %.3f ms  K2 %.3f ms  frac %.4f  e2e %.3f ms" % (j["ms_per_step"], j["roofline"]["kernel_ms_per_launch"], j["roofline"]["frac"], j["e2e"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
  done
done
