#!/bin/bash
# Mechanism check for the intermittent capture failure: an engine finalised by the cyclic collector during a graph capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/flake3.py <<'PY'
import sys, gc, traceback
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import test_gpu_parity as T
from safeopt_b200.swarm import DeviceSwarm
g = T.load_golden("swarm_query_2d")
victims = []
for _ in range(2):
    o = T._swarm_query_problem(g, swarm_backend="device", rng="device", seed=3)
    o.optimize()
    victims.append(o)
torch.cuda.synchronize()
orig = DeviceSwarm._iteration_dev
hits = [0]
def patched(self):
    if torch.cuda.is_current_stream_capturing() and victims:
        victims.pop()                      # a SafeOptSwarm <-> DeviceSwarm cycle becomes garbage inside the capture
        if gc.isenabled():                 # what an automatic collection would do at this point
            hits[0] += 1
            gc.collect()
    orig(self)
DeviceSwarm._iteration_dev = patched
try:
    T.test_device_swarm_device_rng_runs()
    torch.cuda.synchronize()
    print("PASSED; collections inside a capture:", hits[0], flush=True)
except BaseException as e:
    print("FAILED (%s: %s); collections inside a capture: %d" % (type(e).__name__, str(e).splitlines()[0][:120], hits[0]), flush=True)
PY
echo "gc allowed during capture:"; SAFEOPT_B200_GC_IN_CAPTURE=1 timeout 120 python /tmp/flake3.py 2>&1 | tail -2 | cut -c1-250
echo "default:"; timeout 120 python /tmp/flake3.py 2>&1 | tail -2 | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -4 | tee gpurun_out/r2y_pytest_gpu.log
