"""Times so_fit (K1: scale, Ky, blocked Cholesky + inverse, alpha, fragment packing) for a few training-set sizes."""
import json, os, sys, time
os.environ.setdefault("SO_FIT_VERBOSE", "1")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from safeopt_b200.engine import DeviceEngine

out = {}
eng = DeviceEngine(max_gps=1)
for N in (64, 128, 256, 512, 1024, 2048):
    rs = np.random.RandomState(N)
    X = rs.uniform(-2.5, 2.5, (N, 4)); Y = rs.randn(N)
    for _ in range(3):
        eng.fit(0, X, Y, 0, np.ones(4), 2.0, 0.05 ** 2)
    torch.cuda.synchronize()
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.fit(0, X, Y, 0, np.ones(4), 2.0, 0.05 ** 2)     # so_fit synchronises (it reports SO_ERR_NOT_PD)
    torch.cuda.synchronize()
    out["N=%d" % N] = {"fit_ms": (time.perf_counter() - t0) / reps * 1e3}
eng.close()
print(json.dumps(out))
