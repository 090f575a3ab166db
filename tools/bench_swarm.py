"""Config 5 (BASELINE.json): SafeOptSwarm-style particle evaluation at scale -- d=6, 1e5 particles, 2 GPs, N_train=512.
Measures particle evals/s of the device-resident swarm (posterior of all particles for both GPs + fitness + PSO step per
iteration) and, beside it, the oracle port's fitness on the host cores; then one full SafeOptSwarm.optimize() (three swarms of
100 iterations each + the device-side safe-set insertion) through the public class.
Usage: python tools/bench_swarm.py [--iters 20]      (1 GPU)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_swarm.py
The swarm is sharded over the ranks (fixed total: "strong" scaling, as config 5 says "1e5 particles ... 8xB200")."""
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
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = workloads.swarm_workload(args.particles, args.n_train)
    gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                           noise_var=w.noise_var) for i in range(w.n_gps)]
    opt = sb.SafeOptSwarm(gps, [0.0, 0.2], bounds=w.bounds, beta=w.beta, swarm_size=args.particles)
    opt.best_lower_bound = 0.5
    kind = "maximizers"
    opt._fits.refresh()
    swarm = sb.DeviceSwarm(opt._engine, opt.optimal_velocities, lambda p: opt._swarm_fitness(kind, p), bounds=w.bounds, rng="device")
    swarm.init_swarm(w.particles.copy())
    swarm.run_swarm(3)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    swarm.run_swarm(args.iters)
    e1.record()
    sync()
    ms = max_ranks(e0.elapsed_time(e1))
    # fitness only (posterior of both GPs + epilogue), particles resident
    pos = swarm.positions
    e0.record()
    for _ in range(args.iters):
        opt._swarm_fitness(kind, pos)
    e1.record()
    sync()
    ms_fit = max_ranks(e0.elapsed_time(e1))
    # the whole public call: three swarms x (1 + max_iters) fitness evaluations + safe-set re-check + insertion
    np.random.seed(0)
    full = sb.SafeOptSwarm(gps, [0.0, 0.2], bounds=w.bounds, beta=w.beta, swarm_size=args.particles, swarm_backend="device",
                           rng="device")
    full.S = w.X[:64].copy()
    s_before = full.S.shape[0]
    sync()
    t0 = time.perf_counter()
    x_next = full.optimize()
    sync()
    opt_ms = max_ranks(1e3 * (time.perf_counter() - t0))
    launches = full._engine.launches
    t0 = time.perf_counter()
    new = full._select_new_safe_points(full.swarms["expanders"].best_positions, sharded=True)
    sync()
    ins_ms = max_ranks(1e3 * (time.perf_counter() - t0))
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    out = {"config": "C5: swarm d=%d, %d particles, %d GPs, N_train=%d, fp64, %d GPU(s)" % (w.d, args.particles, w.n_gps, args.n_train, world),
           "n_gpus": world, "optimize_ms": opt_ms, "optimize_particle_evals_per_s": 3 * (full.max_iters + 1) * args.particles / (opt_ms * 1e-3),
           "optimize_kernel_launches": launches, "safe_set_before": s_before, "safe_set_after": int(full.S.shape[0]),
           "insertion_ms_on_grown_set": ins_ms, "insertion_accepted_again": int(new.shape[0]), "x_next": [float(v) for v in x_next],
           "pso_iteration_ms": ms / args.iters, "particle_evals_per_s": args.particles * args.iters / (ms * 1e-3),
           "fitness_only_ms": ms_fit / args.iters, "fitness_particle_evals_per_s": args.particles * args.iters / (ms_fit * 1e-3),
           "flops_per_eval": w.n_gps * (args.n_train ** 2 + (3 * w.d + 8) * args.n_train),
           "fitness_tflops": w.n_gps * (args.n_train ** 2 + (3 * w.d + 8) * args.n_train) * args.particles * args.iters / (ms_fit * 1e-3) / 1e12}
    if not args.no_cpu and world == 1:
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
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
