#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in noINV noFACTOR; do
SAFEOPT_B200_LIB=$PWD/tools/ab/$v.so python - 2> gpurun_out/fit_$v.txt <<'PY'
import os, sys
os.environ["SO_FIT_VERBOSE"]="1"
sys.path.insert(0, ".")
import numpy as np, torch
from safeopt_b200.engine import DeviceEngine
eng = DeviceEngine(max_gps=1)
rs = np.random.RandomState(0)
X = rs.uniform(-2.5, 2.5, (256, 4)); Y = rs.randn(256)
for _ in range(4):
    try: eng.fit(0, X, Y, 0, np.ones(4), 2.0, 0.05 ** 2)
    except Exception as e: print("fit raised", str(e)[:80])
PY
echo $v; grep "fit stamps" gpurun_out/fit_$v.txt | tail -1 | cut -c1-500
done
