"""Grid-path posterior kernel time at config 4's grid as the number of observations grows (a BO loop adds one per iteration).
Usage: python tools/time_k2_vs_n.py [N ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import safeopt_b200 as sb
from safeopt_b200 import workloads

ns = [int(a) for a in sys.argv[1:]] or [128, 248, 256, 257, 264, 288, 320, 384, 512]
grid = sb.linearly_spaced_combinations([(-5.0, 5.0)] * 4, 50)
for n in ns:
    w = workloads.grid_workload("t", 4, 50, n)
    gp = sb.GPRegression(w.X, w.Y[:, [0]], kernel=sb.RBF(4, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
    opt = sb.SafeOpt(gp, grid, 0.0, beta=w.beta, threshold=w.threshold)
    for _ in range(2):
        opt.update_confidence_intervals()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.update_confidence_intervals()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    flops = 6.25e6 * (n * n + 20 * n)
    print(json.dumps({"N": n, "k2_ms": round(ms, 3), "canonical_tflops": round(flops / ms * 1e-9, 2)}), flush=True)
    del opt
    torch.cuda.empty_cache()
