"""Config 5 (BASELINE.json): SafeOptSwarm-style particle evaluation at scale -- d=6, 1e5 particles, 2 GPs, N_train=512.
Measures particle evals/s of the device-resident swarm (posterior of all particles for both GPs + fitness + PSO step per
iteration) and, beside it, the oracle port's fitness on the host cores.  Usage: python tools/bench_swarm.py [--iters 20]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=100_000)
    ap.add_argument("--n-train", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import safeopt_b200 as sb
    from safeopt_b200 import workloads
    w = workloads.swarm_workload(args.particles, args.n_train)
    gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                           noise_var=w.noise_var) for i in range(w.n_gps)]
    opt = sb.SafeOptSwarm(gps, [0.0, 0.2], bounds=w.bounds, beta=w.beta, swarm_size=args.particles)
    opt.best_lower_bound = 0.5
    kind = "maximizers"
    swarm = sb.DeviceSwarm(opt._engine, opt.optimal_velocities, lambda p: opt._fitness_device(kind, p), bounds=w.bounds, rng="device")
    swarm.init_swarm(w.particles.copy())
    swarm.run_swarm(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    swarm.run_swarm(args.iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # fitness only (posterior of both GPs + epilogue), particles resident
    pos = swarm.positions
    e0.record()
    for _ in range(args.iters):
        opt._fitness_device(kind, pos)
    e1.record()
    torch.cuda.synchronize()
    ms_fit = e0.elapsed_time(e1)
    out = {"config": "C5: swarm d=%d, %d particles, %d GPs, N_train=%d, fp64, 1 GPU" % (w.d, args.particles, w.n_gps, args.n_train),
           "pso_iteration_ms": ms / args.iters, "particle_evals_per_s": args.particles * args.iters / (ms * 1e-3),
           "fitness_only_ms": ms_fit / args.iters, "fitness_particle_evals_per_s": args.particles * args.iters / (ms_fit * 1e-3),
           "flops_per_eval": w.n_gps * (args.n_train ** 2 + (3 * w.d + 8) * args.n_train),
           "fitness_tflops": w.n_gps * (args.n_train ** 2 + (3 * w.d + 8) * args.n_train) * args.particles * args.iters / (ms_fit * 1e-3) / 1e12}
    if not args.no_cpu:
        from oracle import gpy_lite, safeopt_port as port
        go = [gpy_lite.GPRegression(w.X, w.Y[:, [i]], kernel=gpy_lite.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                                    noise_var=w.noise_var) for i in range(w.n_gps)]
        t0 = time.perf_counter()
        vo, so = port.particle_fitness(go, np.array([0.0, 0.2]), w.beta, opt.scaling, kind, w.particles, best_lower_bound=0.5)
        dt = time.perf_counter() - t0
        vd, sd = opt._compute_particle_fitness(kind, w.particles)
        out["cpu_port_particle_evals_per_s"] = args.particles / dt
        out["cpu_cores"] = os.cpu_count()
        out["max_abs_diff_vs_port"] = float(np.abs(vd - vo).max())
        out["safe_mismatch"] = int((sd != so).sum())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
