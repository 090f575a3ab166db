"""cProfile of the host side of SafeOpt.optimize() on a small configuration (where the device work is a few tens of
microseconds and the step is host-bound).  Usage: python tools/profile_host.py [C2|C3]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import safeopt_b200 as sb
from safeopt_b200 import workloads

w = workloads.config(sys.argv[1] if len(sys.argv) > 1 else "C2")
grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
       for i in range(w.n_gps)]
opt = sb.SafeOpt(gps if w.n_gps > 1 else gps[0], grid, w.fmin if w.n_gps > 1 else w.fmin[0], beta=w.beta, threshold=w.threshold)
for _ in range(20):
    opt.optimize()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(500):
    opt.optimize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
