"""Time the posterior kernel alone (config 4 by default) for every library variant given on the command line
(tools/ab/<name>.so, built by tools/build_variant.sh); results may be garbage for the timing-diagnostic variants, only the
duration is read.  Usage: python tools/time_k2.py [name ...]; each variant runs in a subprocess (one library per process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import safeopt_b200 as sb
    from safeopt_b200 import workloads
    w = workloads.config(os.environ.get("K2_CONFIG", "C4"))
    grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
    gp = sb.GPRegression(w.X, w.Y[:, [0]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
    opt = sb.SafeOpt(gp, grid, 0.0, beta=w.beta, threshold=w.threshold)
    for _ in range(3):
        opt.update_confidence_intervals()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.update_confidence_intervals()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(json.dumps({"k2_ms_median": float(np.median(ts)), "k2_ms_min": float(np.min(ts))}))


if __name__ == "__main__":
    if os.environ.get("K2_CHILD"):
        child()
    else:
        names = sys.argv[1:] or sorted(f[:-3] for f in os.listdir(os.path.join(ROOT, "tools", "ab")) if f.endswith(".so"))
        for rep in range(int(os.environ.get("K2_ROUNDS", "2"))):
            for n in names:
                env = dict(os.environ, K2_CHILD="1", SAFEOPT_B200_LIB=os.path.join(ROOT, "tools", "ab", n + ".so"))
                r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
                print(n, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
