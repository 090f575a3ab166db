#!/usr/bin/env python
"""Benchmark of the SafeOpt hot path (BASELINE.json metric: grid-point GP posterior + safe-set evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Grid configs (C1..C4): a *step* is one ``SafeOpt.optimize()`` over the whole candidate grid: GP posterior of every row
(kernel rows, L^-1 contraction, mean/var), confidence bounds, safe set, maximisers, expander candidates / search and the
query-point argmax.  Default workload = BASELINE config 4 (the configuration the metric is quoted on): d=4, 50^4 = 6.25e6
rows, N_train=256, 1 GP (objective = constraint), fp64, synthetic RBF problem of SURVEY.md section 8d.  With N ranks the
FIXED 50^4 grid is split into N contiguous row blocks (``--scaling strong``, the default: BASELINE's "50^4 ... sharded
8xB200"); ``--scaling weak`` keeps a full config-4 block per rank instead (the grid becomes 50 x (50 N) x 50 x 50).
Config C5 (``--config C5``): SafeOptSwarm, d=6, 1e5 particles, 2 GPs, N_train=512; a step is one PSO iteration of the
device-resident swarm (update, posterior of every particle for both GPs, fitness, bests, global best); the particles are
split over the ranks.  The default (C4) line also carries a short C5 measurement under ``secondary``.

Reported: ``value`` (device-resident throughput, fit cached), ``e2e`` (through the public API with host inputs),
``roofline`` of the dominant kernel (fp64 tensor pipe; peak = cuBLAS DGEMM measured in this session, because
MEASURED_PEAKS.json only carries HBM and bf16 numbers), ``cpu_baseline`` (oracle port on the host cores on a bounded
sample of the same workload), ``parity`` (this run's masks / bounds against the port on that sample; with N > 1 the
sharded answer against a single-GPU evaluation of the same grid), ``sharded_parity`` (reference-generated golden
fixtures re-run under this run's communicator) and the clocks seen during timing.  A parity mismatch exits non-zero.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid_point_posterior_safe_set_evals_per_sec"
UNIT = "evals/s"
DMMA_MICROBENCH_TFLOPS = 37.1          # profiles/r01_fp64_rates_b200.jsonl: DMMA.8x8x4 issue-bound rate on this pool's B200


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--num-samples", type=int, default=None, help="override points per axis (development only)")
    ap.add_argument("--particles", type=int, default=None, help="override the C5 swarm size (development only)")
    ap.add_argument("--cpu-sample-rows", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short C5 measurement inside the C4 line")
    ap.add_argument("--no-sharded-parity", action="store_true", help="skip the golden fixtures under the communicator")
    ap.add_argument("--explicit-rows", action="store_true", help="force the explicit-rows kernel path (no grid tables)")
    ap.add_argument("--fp64", action="store_true", help="C3: run the fp64 path instead of the fp32 mode the config names")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong: the config's grid is split N ways (default); weak: per-GPU rows fixed (grid axis 1 grows with N)")
    return ap.parse_args()


def flops_per_eval(n_train, d, n_gps):
    """Canonical algorithmic work per grid row: G * (N^2 + (3d+8) N)."""
    return n_gps * (n_train * n_train + (3 * d + 8) * n_train)


def bytes_per_eval(d, n_gps, grid_path):
    """Algorithmic HBM bytes per row: candidate in (0 when rows are generated), l/u out, S/M bytes."""
    return (0 if grid_path else d * 8) + 2 * n_gps * 8 + 3


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    smax = float(parts[1])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            hi = [c for c in sm if c >= 0.5 * max(sm)]
            out.update(sm_mhz=float(np.median(hi)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_port_run(w, sample_rows, steps=1, warmup=0):
    """Time the oracle port (NumPy/SciPy restatement of the reference path) on a bounded row sample.
    Returns (first row of the sample, rows, per-step seconds, the port problem with its Q / S of the last step)."""
    from oracle import gpy_lite, safeopt_port as port
    M = w.n_rows
    take = min(sample_rows, M)
    # contiguous block around the training data (the block contains safe, unsafe and boundary rows)
    start = max(0, M // 2 - take // 2)
    sub = port.grid_rows(w.bounds, w.num_samples, np.arange(start, start + take))
    gps = [gpy_lite.GPRegression(w.X, w.Y[:, [i]], kernel=gpy_lite.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                                 noise_var=w.noise_var) for i in range(w.n_gps)]
    prob = port.GridProblem.create(gps, sub, w.fmin, beta=w.beta, threshold=w.threshold)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        try:
            prob.optimize(chunk=100_000)
        except port.NoSafePoints:       # a small sample block may hold no safe row; the posterior + set passes ran
            pass
        times.append(time.perf_counter() - t0)
    times = times[warmup:]
    return start, take, times, prob


def cpu_swarm_run(w, sample, steps=1, warmup=0):
    """C5 on the host cores: the port's particle fitness (posterior of both GPs + fitness epilogue) on a particle sample."""
    from oracle import gpy_lite, safeopt_port as port
    gps = [gpy_lite.GPRegression(w.X, w.Y[:, [i]], kernel=gpy_lite.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                                 noise_var=w.noise_var) for i in range(w.n_gps)]
    take = min(sample, w.n_particles)
    pts = w.particles[:take]
    scaling = np.sqrt(np.full(w.n_gps, w.variance))
    times, out = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = port.particle_fitness(gps, np.asarray(SWARM_FMIN), w.beta, scaling, "maximizers", pts, best_lower_bound=0.5)
        times.append(time.perf_counter() - t0)
    return take, times[warmup:], out


SWARM_FMIN = [0.0, 0.2]


def run_reference(args, w, rank, world):
    """--impl reference: the reference algorithm's CPU path (oracle port) on the host cores."""
    if rank != 0:
        return
    cores = os.cpu_count()
    if args.config == "C5":
        take, times, _ = cpu_swarm_run(w, 20_000, steps=args.steps, warmup=args.warmup)
        sample = ("%d of the %d particles per step (oracle/safeopt_port.particle_fitness over oracle/gpy_lite.py, NumPy/OpenBLAS "
                  "threads=%d)" % (take, w.n_particles, cores))
        config = swarm_config(w, args)
    else:
        _, take, times, _ = cpu_port_run(w, args.cpu_sample_rows, steps=args.steps, warmup=args.warmup)
        sample = ("%d contiguous grid rows of the %d-row workload per step (oracle/safeopt_port.py over oracle/gpy_lite.py, "
                  "NumPy/OpenBLAS threads=%d, 100k-row chunks); rows/s does not depend on the block size, so the whole-grid "
                  "figure is this rate" % (take, w.n_rows, cores))
        config = workload_config(w, args, grid_path=None)
    total = float(np.sum(times))
    value = take * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(w, args, grid_path, dtype_label=None):
    return {"workload": "%s: %dD RBF-ARD, %d constraint GP(s), %s=%d grid rows, N_train=%d, %s, beta=%g, fmin=0, threshold=%g" % (
        w.name, w.d, w.n_gps, "x".join(str(n) for n in w.samples_per_axis), w.n_rows, w.n_train, dtype_label or w.dtype, w.beta,
        w.threshold),
        "rows": w.n_rows, "rows_per_gpu": -(-w.n_rows // args.gpus), "n_train": w.n_train, "d": w.d, "n_gps": w.n_gps,
        "parallelism": "rows sharded x%d (%s scaling)" % (args.gpus, args.scaling),
        "candidate_path": None if grid_path is None else ("grid rows generated on device" if grid_path else "explicit rows in HBM"),
        "l2": "each step streams 33 B/row of outputs (206 MB at C4, > 126 MB L2) and re-derives everything else; "
              "operands (L^-1, tables, 0.8 MB) are meant to stay cache-resident"}


def swarm_config(w, args):
    return {"workload": "C5: SafeOptSwarm %dD RBF-ARD, %d particles, %d GPs (shared factorisation), N_train=%d, fp64, beta=%g" % (
        w.d, w.n_particles, w.n_gps, w.n_train, w.beta),
        "rows": w.n_particles, "rows_per_gpu": -(-w.n_particles // args.gpus), "n_train": w.n_train, "d": w.d, "n_gps": w.n_gps,
        "parallelism": "particles sharded x%d (strong scaling)" % args.gpus, "candidate_path": "explicit rows in HBM",
        "l2": "every iteration rewrites positions/velocities (9.6 MB at 1e5 x 6) and re-derives every kernel row; the operands "
              "(L^-1 packed, 1.05 MB at N=512) stay L2-resident"}


# ----------------------------------------------------------------------------- helpers of the B200 arm
class Ctx:
    def __init__(self, torch, dev, dist, rank, world):
        self.torch, self.dev, self.dist, self.rank, self.world = torch, dev, dist, rank, world

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok):
        if self.dist is None:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() == 1)


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)


def _unpack(packed, n):
    return np.unpackbits(packed)[:n].astype(bool)


def sharded_parity(ctx):
    """Reference-generated golden fixtures (tests/golden, made by oracle/make_golden.py from the unmodified reference)
    re-run under THIS run's communicator: grid cases with the GP and the Lipschitz expander rule, and a whole
    SafeOptSwarm.optimize() on the sharded device swarm following the reference's random stream."""
    import safeopt_b200 as sb
    kern_cls = {0: sb.RBF, 1: sb.Matern32, 2: sb.Matern52}
    failed, cases, worst_dq = [], 0, 0.0
    for name in ["expander_g2", "lipschitz_g2", "config_C4_n10"]:
        g = _golden(name)
        X, Y, d = g["X"], g["Y"], g["X"].shape[1]
        grid = sb.linearly_spaced_combinations([tuple(b) for b in g["bounds"]], [int(v) for v in np.atleast_1d(g["num_samples"])])
        gps = [sb.GPRegression(X, Y[:, [i]], kernel=kern_cls[int(g["kind"])](d, variance=float(g["variance"]),
                                                                             lengthscale=np.asarray(g["lengthscale"], dtype=float), ARD=True),
                               noise_var=float(g["noise_var"]), device=ctx.dev) for i in range(Y.shape[1])]
        lip = np.atleast_1d(g["lipschitz"]) if "lipschitz" in g.files else np.zeros(0)
        lip = None if lip.size == 0 else ([float(v) for v in lip] if lip.size > 1 else float(lip[0]))
        fmin = [float(v) for v in g["fmin"]]
        opt = sb.SafeOpt(gps if len(gps) > 1 else gps[0], grid, fmin if len(gps) > 1 else fmin[0], lipschitz=lip,
                         beta=float(g["beta"]), threshold=float(g["threshold"]), device=ctx.dev)
        x = opt.optimize()
        n = int(g["n_rows"])
        dq = float(np.abs(opt.Q - g["Q"]).max())
        worst_dq = max(worst_dq, dq)
        ok = (opt.last_query_row == int(g["row_next"]) and np.array_equal(x, g["x_next"]) and dq < 1e-8
              and np.array_equal(opt.S, _unpack(g["S"], n)) and np.array_equal(opt.M, _unpack(g["M"], n))
              and np.array_equal(opt.G, _unpack(g["G"], n)))
        cases += 1
        if not ctx.all_true(ok):
            failed.append(name)
    g = _golden("swarm_query_2d")
    Xs, Ys, ds = g["X"], g["Y"], g["X"].shape[1]
    sgps = [sb.GPRegression(Xs, Ys[:, [i]], kernel=kern_cls[int(g["kind"])](ds, variance=float(g["variance"]),
                                                                            lengthscale=np.asarray(g["lengthscale"], dtype=float), ARD=True),
                            noise_var=float(g["noise_var"]), device=ctx.dev) for i in range(Ys.shape[1])]
    opt = sb.SafeOptSwarm(sgps, list(g["fmin"]), bounds=[tuple(b) for b in g["bounds"]], beta=float(g["beta"]),
                          swarm_size=int(g["swarm_size"]), device=ctx.dev, swarm_backend="device", rng="host")
    opt.max_iters = int(g["max_iters"])
    np.random.seed(int(g["seed"]))
    x = opt.optimize()
    ok = (np.abs(x - g["x_next"]).max() < 1e-7 and opt.S.shape == g["S_final"].shape and np.abs(opt.S - g["S_final"]).max() < 1e-7)
    cases += 1
    if not ctx.all_true(ok):
        failed.append("swarm_query_2d")
    return {"cases": cases, "ok": not failed, "failed": failed, "max_abs_dQ": worst_dq, "world": ctx.world,
            "what": "tests/golden fixtures generated by the unmodified reference, re-run sharded over this run's ranks: Q within "
                    "1e-8, S/M/G masks, query row and point identical; swarm trajectory within 1e-7"}


def measure_dgemm_tflops(torch, dev, n=8192):
    """cuBLAS DGEMM rate in this session = the fp64 roofline denominator (burst, timed alone, best of 10 after 2 warm-ups)."""
    try:
        a = torch.zeros((n, n), dtype=torch.float64, device=dev)
        b = torch.zeros((n, n), dtype=torch.float64, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        del a, b
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "torch.matmul fp64 %d^3 (cuBLAS DGEMM), best of 10, this session" % n
    except Exception as exc:  # pragma: no cover
        return 35.76, "fallback: profiles/r01_fp64_rates_b200.jsonl cublas_dgemm n=8192 (%s)" % exc


def time_k2(eng, torch):
    """Wrap the engine's posterior launches with CUDA events (on the launching stream); returns (events list, restore())."""
    events = []
    names = ["posterior_grid", "posterior_rows", "posterior_multi"] + (["posterior_grid_f32"] if hasattr(eng, "posterior_grid_f32") else [])
    orig = {n: getattr(eng, n) for n in names}

    def timed(fn):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            events.append((e0, e1))
            return out
        return wrapper

    for n in names:
        setattr(eng, n, timed(orig[n]))
    # unchanged inputs: update_confidence_intervals re-issues its remembered launches through replay_tape (one call per step)
    if hasattr(eng, "replay_tape"):
        names.append("replay_tape")
        orig["replay_tape"] = eng.replay_tape
        eng.replay_tape = timed(orig["replay_tape"])

    def restore():
        for n in names:
            setattr(eng, n, orig[n])
    return events, restore


def hbm_peak():
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            return json.load(open(f)).get("hbm_gbs")
        except Exception:
            pass
    return None


# ----------------------------------------------------------------------------- B200 arm: grid configs
def run_grid(args, w, ctx, sampler_index):
    import safeopt_b200 as sb
    torch, dev, rank, world = ctx.torch, ctx.dev, ctx.rank, ctx.world

    grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
    if args.explicit_rows:
        os.environ["SAFEOPT_B200_GRID_FAST_PATH"] = "0"      # explicit-rows kernels (200 MB of candidates in HBM)
    fp32 = (w.dtype == "fp32") and not args.fp64 and hasattr(sb.SafeOpt, "precision")

    def make(distributed=True):
        gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                               noise_var=w.noise_var, device=dev) for i in range(w.n_gps)]
        kw = {"precision": "fp32"} if fp32 else {}
        return sb.SafeOpt(gps if w.n_gps > 1 else gps[0], grid, w.fmin if w.n_gps > 1 else w.fmin[0], beta=w.beta,
                          threshold=w.threshold, device=dev, distributed=distributed, **kw)

    opt = make()
    grid_path = opt._grid_axes is not None
    eng = opt._engine

    # ---- device-resident throughput: fit cached, everything else per step
    for _ in range(args.warmup):
        x_next = opt.optimize()
    ctx.barrier()
    sampler = ClockSampler(sampler_index)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    k2_events, restore = time_k2(eng, torch)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    for _ in range(args.steps):
        x_next = opt.optimize()
    ev1.record()
    ctx.barrier()
    restore()
    dev_ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launches - launches0
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in k2_events]))
    n_k2 = len(k2_events)
    clocks = sampler.stop() if rank == 0 else None
    n_safe, n_max = opt._safe_info["n_safe"], opt._max_info["n_max"] if opt._max_info else 0
    trace = dict(opt.last_trace)
    row_next = int(opt.last_query_row)

    # ---- end to end through the public API with host inputs: refit from host X/Y + optimize + result to host
    X_pin = torch.from_numpy(np.ascontiguousarray(w.X)).pin_memory()
    Y_pin = torch.from_numpy(np.ascontiguousarray(w.Y)).pin_memory()
    h2d = X_pin.numel() * 8 + Y_pin.numel() * 8 + (w.d + 2) * 8 * w.n_gps
    d2h = 136 * world + 16 + w.n_gps * 4    # every rank's two 64-byte records + candidate count (one copy), fit status words

    def e2e_step():
        Xh, Yh = X_pin.numpy(), Y_pin.numpy()
        for i, gp in enumerate(opt.gps):
            gp.set_XY(Xh, Yh[:, [i]])
        opt._fits.invalidate()          # a new observation would change the fingerprint; force the refit
        return np.asarray(opt.optimize())

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    ctx.barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        x_e2e = e2e_step()
    ev1.record()
    ctx.barrier()
    e2e_ms = ctx.max_over_ranks(max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0)))

    rows = w.n_rows
    value = rows * args.steps / (dev_ms * 1e-3)
    e2e_value = rows * args.steps / (e2e_ms * 1e-3)
    local_rows = opt._row1 - opt._row0

    # ---- parity of THIS run's answer
    parity = {}
    if world > 1:
        # the sharded answer against a single-GPU evaluation of the same grid, on rank 0
        single = None
        if rank == 0:
            ref = make(distributed=False)
            x_ref = ref.optimize()
            single = {"x_next": [float(v) for v in x_ref], "row_next": int(ref.last_query_row), "n_safe": int(ref._safe_info["n_safe"]),
                      "n_maximizers": int(ref._max_info["n_max"]) if ref._max_info else 0,
                      "n_expander_candidates": int(ref.last_trace.get("n_candidates", 0))}
            same = (np.array_equal(x_ref, x_next) and single["row_next"] == row_next and single["n_safe"] == int(n_safe)
                    and single["n_maximizers"] == int(n_max) and single["n_expander_candidates"] == int(trace.get("n_candidates", 0)))
            parity["vs_single_gpu"] = {"equal": bool(same), "single_gpu": single}
            del ref
        ctx.barrier()
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        start, take, times, prob = cpu_port_run(w, args.cpu_sample_rows, steps=1, warmup=0)
        cpu = {"value": take / times[0], "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "%d contiguous grid rows of the %d-row workload, one optimize() (oracle/safeopt_port.py over "
                         "oracle/gpy_lite.py, NumPy/OpenBLAS default threads, 100k-row chunks), %.1f s; rows/s does not depend on "
                         "the block size" % (take, rows, times[0])}
        # the port's Q / S on its block against this run's; M through the device's global max l0[S] (the port saw only the block)
        Qd = opt._Q_d[start:start + take].cpu().numpy()
        Sd = opt._S_d[start:start + take].cpu().numpy().astype(bool)
        Md = opt._M_d[start:start + take].cpu().numpy().astype(bool)
        max_l0 = opt._safe_info["max_l0"]
        fm = np.asarray(w.fmin, dtype=float)
        tol = (3e-4 if fp32 else 1e-9) * 2 * np.sqrt(w.variance)
        dq = float(np.abs(Qd - prob.Q).max())
        margin_s = float(np.abs(prob.Q[:, ::2] - fm).min())
        margin_m = float(np.abs(prob.Q[prob.S, 1] - max_l0).min()) if prob.S.any() else float("inf")
        Mp = prob.S & (prob.Q[:, 1] >= max_l0)
        extra = {}
        if fp32:
            # fp32 arithmetic (north star: posterior mean / var within 1e-4 relative): mean and variance recovered from the
            # bounds, l/u = mean -/+ beta sd; masks are compared outside the band the bound error can reach around the thresholds
            mean_d, mean_p = 0.5 * (Qd[:, ::2] + Qd[:, 1::2]), 0.5 * (prob.Q[:, ::2] + prob.Q[:, 1::2])
            var_d, var_p = ((Qd[:, 1::2] - Qd[:, ::2]) / (2 * w.beta)) ** 2, ((prob.Q[:, 1::2] - prob.Q[:, ::2]) / (2 * w.beta)) ** 2
            dmean, dvar = float(np.abs(mean_d - mean_p).max()), float(np.abs(var_d - var_p).max())
            tol_mean, tol_var = 1e-4 * max(1.0, float(np.abs(w.Y).max())), 1e-4 * w.variance
            band_s = np.all(np.abs(prob.Q[:, ::2] - fm) > tol, axis=1)
            band_m = band_s & (np.abs(prob.Q[:, 1] - max_l0) > 2 * tol)
            masks_equal = bool(np.array_equal(Sd[band_s], prob.S[band_s]) and np.array_equal(Md[band_m], Mp[band_m]))
            extra = {"max_abs_dmean": dmean, "tolerance_dmean": tol_mean, "max_abs_dvar": dvar, "tolerance_dvar": tol_var,
                     "rows_inside_band": int((~band_m).sum()), "mask_rows_differing_inside_band": int((Sd != prob.S).sum() + (Md != Mp).sum())}
            ok = bool(masks_equal and dmean <= tol_mean and dvar <= tol_var)
        else:
            masks_equal = bool(np.array_equal(Sd, prob.S) and np.array_equal(Md, Mp))
            ok = bool(masks_equal and dq < tol)
        parity["vs_port"] = dict({"rows": int(take), "first_row": int(start), "masks_equal": masks_equal, "max_abs_dQ": dq, "tolerance_dQ": tol,
                                  "min_margin_S": margin_s, "min_margin_M": margin_m, "n_safe_in_block": int(prob.S.sum()),
                                  "n_maximizers_in_block": int(Mp.sum()), "ok": ok}, **extra)

    # ---- roofline of the dominant kernel: fp64 tensor pipe (or the tf32 tensor pipe in fp32 mode)
    fpe = flops_per_eval(w.n_train, w.d, 1)
    achieved = fpe * local_rows / (k2_ms * 1e-3) / 1e12
    line = None
    if rank == 0:
        if fp32:
            peak, peak_src = measure_tf32_tflops(torch, dev)
            peak, peak_src = peak / 3.0, peak_src + "; / 3 (error-compensated 3xTF32: three tensor-core products per algorithmic product)"
            kname = "k_posterior_f32 (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators)"
        else:
            peak, peak_src = measure_dgemm_tflops(torch, dev)
            kname = "k_posterior (DMMA.8x8x4 fp64 tensor pipe)"
        traffic = None
        tf = os.path.join(ROOT, "profiles", "k_posterior_traffic.json")
        if os.path.exists(tf) and args.config == "C4" and not fp32:
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
                if traffic is not None and world > 1:
                    traffic = traffic / world        # the capture is the single-GPU launch; a rank of R reads 1/R of the A' table
            except Exception:
                pass
        per_step = n_k2 / args.steps
        roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src, "flops_per_eval": fpe,
                    "evals_per_launch": local_rows, "kernel_ms_per_launch": k2_ms, "kernel_launches_per_step": per_step,
                    "kernel_share_of_step": k2_ms * per_step / (dev_ms / args.steps),
                    "algorithmic_hbm_gbs": bytes_per_eval(w.d, w.n_gps, grid_path) * local_rows / (k2_ms * 1e-3) / 1e9,
                    "hbm_peak_gbs_measured": hbm_peak(),
                    "note": "MEASURED_PEAKS.json has no fp64 / tf32 number, so the denominator is the matching cuBLAS GEMM measured in this session"}
        if not fp32:
            roofline["fp64_dmma_microbench_tflops"] = DMMA_MICROBENCH_TFLOPS
            roofline["frac_of_dmma_microbench"] = achieved / DMMA_MICROBENCH_TFLOPS
            roofline["traffic_note"] = ("the measured DRAM traffic is dominated by the precomputed scaled-operand table A'(s) (270 KB per "
                                        "slow index), read exactly once per launch; it replaces N fp64 exp per row on the FP64 pipe the "
                                        "contraction saturates. Outputs: 33 B/row")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32" if fp32 else "f64", "data": "synthetic",
            "config": workload_config(w, args, grid_path, "fp32 (3xTF32 tensor cores; fit and set logic fp64)" if fp32 else "fp64"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "set_XY from pinned host X/Y -> device refit (Cholesky, L^-1, tables) -> optimize() -> next parameters on host"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "exchange": "records written into every rank's buffer over NVLink by the set-pass kernel itself (CUDA IPC peer mapping)"
                        if (world > 1 and opt._peer) else ("torch.distributed all-gathers between chained kernels" if world > 1 else "single rank"),
            "result": {"x_next": [float(v) for v in np.asarray(x_next)], "row_next": row_next, "n_safe": int(n_safe), "n_maximizers": int(n_max),
                       "n_expander_candidates": int(trace.get("n_candidates", 0)), "e2e_same_x": bool(np.array_equal(x_next, x_e2e))},
        }
    del opt
    return line


def measure_tf32_tflops(torch, dev, n=8192):
    """cuBLAS TF32 GEMM rate in this session (fp32 storage, tf32 tensor cores), best of 10."""
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.zeros((n, n), dtype=torch.float32, device=dev)
        b = torch.zeros((n, n), dtype=torch.float32, device=dev)
        for _ in range(2):
            torch.matmul(a, b)
        torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "torch.matmul fp32/tf32 %d^3 (cuBLAS), best of 10, this session" % n
    except Exception as exc:  # pragma: no cover
        return 1100.0, "fallback: nominal dense TF32 (%s)" % exc


# ----------------------------------------------------------------------------- B200 arm: config 5 (swarm)
def run_swarm(args, w, ctx, sampler_index, steps, warmup, with_e2e=True, with_cpu=True):
    """C5: PSO iterations of the device-resident swarm (value) and whole SafeOptSwarm.optimize() calls from host data (e2e)."""
    import safeopt_b200 as sb
    torch, dev, rank, world = ctx.torch, ctx.dev, ctx.rank, ctx.world
    P = w.n_particles

    def gps():
        return [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                                noise_var=w.noise_var, device=dev) for i in range(w.n_gps)]

    opt = sb.SafeOptSwarm(gps(), list(SWARM_FMIN), bounds=w.bounds, beta=w.beta, swarm_size=P, device=dev, swarm_backend="device",
                          rng="device")
    opt.best_lower_bound = 0.5
    opt._fits.refresh()
    eng = opt._engine
    swarm = opt.swarms["maximizers"]
    swarm.init_swarm(w.particles.copy())
    swarm.run_swarm(max(warmup, 3))
    ctx.barrier()
    sampler = ClockSampler(sampler_index)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    swarm.run_swarm(steps)
    ev1.record()
    ctx.barrier()
    dev_ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    gbest, gval = swarm.global_best, swarm.global_best_value

    # the dominant kernel alone: posterior of this rank's particles for both GPs (one contraction), CUDA events on the stream
    pos = swarm.positions
    k2 = []
    for _ in range(3):
        opt._swarm_fitness("maximizers", pos)
    for _ in range(max(steps, 5)):
        mean, var, values, safe = opt._fitness_buffers(pos.shape[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.posterior_multi(list(range(w.n_gps)), pos, 0, pos.shape[0], w.beta, [-np.inf] * w.n_gps,
                            means=[mean[i] for i in range(w.n_gps)], variances=[var[i] for i in range(w.n_gps)])
        e1.record()
        k2.append((e0, e1))
    torch.cuda.synchronize(dev)
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in k2]))

    # parity: this rank's particles against the port (bounded sample, rank 0)
    parity = None
    cpu = None
    if rank == 0 and with_cpu and not args.no_cpu_baseline:
        take, times, (vo, so) = cpu_swarm_run(w, 20_000 if world == 1 else 4_000)
        vd, sd = opt._compute_particle_fitness("maximizers", w.particles[:take])
        parity = {"vs_port": {"particles": int(take), "max_abs_dvalue": float(np.abs(vd - vo).max()), "safe_equal": bool(np.array_equal(sd, so)),
                              "ok": bool(np.array_equal(sd, so) and np.abs(vd - vo).max() < 1e-8)}}
        if world == 1:
            cpu = {"value": take / times[0], "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": "%d of the %d particles, one fitness pass (oracle/safeopt_port.particle_fitness over oracle/gpy_lite.py, "
                             "NumPy/OpenBLAS default threads), %.1f s" % (take, P, times[0])}
    if world > 1 and with_cpu:
        ctx.barrier()

    e2e = None
    if with_e2e:
        # the public call from host data: three swarms x (1 + max_iters) fitness passes + safe-set maintenance; the particles are
        # sampled from the host-resident safe set and copied to the device inside the call
        np.random.seed(0)
        full = sb.SafeOptSwarm(gps(), list(SWARM_FMIN), bounds=w.bounds, beta=w.beta, swarm_size=P, device=dev, swarm_backend="device",
                               rng="device")
        X_pin = torch.from_numpy(np.ascontiguousarray(w.X)).pin_memory()
        Y_pin = torch.from_numpy(np.ascontiguousarray(w.Y)).pin_memory()
        n_calls = 2
        times = []
        for it in range(1 + n_calls):
            full.S = w.X[:64].copy()
            full.best_lower_bound = -np.inf
            full.greedy_point = full.S[0, :]
            for i, gp in enumerate(full.gps):
                gp.set_XY(X_pin.numpy(), Y_pin.numpy()[:, [i]])
            full._fits.invalidate()
            ctx.barrier()
            t0 = time.perf_counter()
            x_next = full.optimize()
            torch.cuda.synchronize(dev)
            times.append(1e3 * (time.perf_counter() - t0))
        opt_ms = ctx.max_over_ranks(float(np.mean(times[1:])))
        evals = 3 * (full.max_iters + 1) * P
        e2e = {"value": evals / (opt_ms * 1e-3), "unit": UNIT, "ms_per_step": opt_ms,
               "h2d_bytes_per_step": int(X_pin.numel() * 8 + Y_pin.numel() * 8 + 3 * P * w.d * 8 // world),
               "d2h_bytes_per_step": int(3 * (w.d + 4) * 8 + full.S.shape[0] * w.d * 8),
               "what": "SafeOptSwarm.optimize() from host data: set_XY + refit, three swarms of %d particles x (1 + %d) fitness passes, "
                       "safe-set re-check and correlation-filtered insertion (|S| 64 -> %d), next parameters on host; "
                       "evals = 3 x %d x particles" % (P, full.max_iters, full.S.shape[0], full.max_iters + 1),
               "x_next": [float(v) for v in x_next]}
        del full

    if rank != 0:
        return None
    fpe = flops_per_eval(w.n_train, w.d, w.n_gps)
    fpe_exec = w.n_train * w.n_train + (3 * w.d + 8) * w.n_train + (w.n_gps - 1) * 2 * w.n_train   # one shared contraction + one V.z per further GP
    local = pos.shape[0]
    peak, peak_src = measure_dgemm_tflops(torch, dev)
    achieved = fpe_exec * local / (k2_ms * 1e-3) / 1e12
    return {
        "metric": METRIC, "value": P * steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": swarm_config(w, args), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "k_posterior (explicit rows: fp64 kernel-row generation + DMMA.8x8x4 contraction, two GPs "
                                                   "sharing one contraction)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": None, "peak_source": peak_src,
                     "flops_per_eval": fpe_exec, "flops_per_eval_canonical_two_contractions": fpe, "evals_per_launch": local,
                     "kernel_ms_per_launch": k2_ms, "kernel_launches_per_step": 1, "kernel_share_of_step": k2_ms / (dev_ms / steps),
                     "fp64_dmma_microbench_tflops": DMMA_MICROBENCH_TFLOPS, "frac_of_dmma_microbench": achieved / DMMA_MICROBENCH_TFLOPS,
                     "note": "EXECUTED flops: the two GPs share X, kernel and noise, so one contraction serves both (the canonical count "
                             "of SURVEY 8d assumes one contraction per GP and would read 2x this)"},
        "cpu_baseline": cpu, "parity": parity,
        "exchange": ("best records written into every rank's buffer over NVLink by the update kernel itself" if swarm.in_kernel_exchange
                     else "one all-gather per iteration") if world > 1 else "single rank",
        "graph": bool(swarm.use_graph),
        "result": {"global_best": [float(v) for v in gbest], "global_best_value": float(gval)},
    }


# ----------------------------------------------------------------------------- B200 arm driver
def run_b200(args, w, rank, world, local_rank):
    import torch
    from safeopt_b200 import workloads
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
    ctx = Ctx(torch, dev, dist, rank, world)

    sp = None
    if not args.no_sharded_parity:
        sp = sharded_parity(ctx)
        ctx.barrier()

    if args.config == "C5":
        line = run_swarm(args, w, ctx, local_rank, args.steps, args.warmup)
    else:
        line = run_grid(args, w, ctx, local_rank)
        if args.config == "C4" and not args.no_secondary and args.num_samples is None:
            ws = workloads.swarm_workload(args.particles or 100_000)
            sec = run_swarm(args, ws, ctx, local_rank, steps=20, warmup=3, with_e2e=True, with_cpu=(world == 1))
            if rank == 0:
                keep = ("value", "unit", "ms_per_step", "steps", "scaling", "config", "e2e", "gpu_launches", "roofline", "parity", "exchange",
                        "graph", "cpu_baseline")
                line["secondary"] = [dict({"name": "C5"}, **{k: sec[k] for k in keep})]
    if rank != 0:
        return 0
    line["sharded_parity"] = sp
    print(json.dumps(line), flush=True)
    bad = []
    if sp is not None and not sp["ok"]:
        bad.append("sharded_parity: %s" % sp["failed"])
    for src in [line] + line.get("secondary", []):
        par = src.get("parity") or {}
        if "vs_port" in par and not par["vs_port"]["ok"]:
            bad.append("parity.vs_port")
        if "vs_single_gpu" in par and not par["vs_single_gpu"]["equal"]:
            bad.append("parity.vs_single_gpu")
    if bad:
        sys.stderr.write("bench.py: PARITY MISMATCH: %s\n" % "; ".join(bad))
        return 3
    return 0


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from safeopt_b200 import workloads
    if args.impl == "b200" and args.gpus != world:
        # one process per GPU: --gpus N is only meaningful under torchrun with N ranks
        sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE=%d; running with %d rank(s)\n" % (args.gpus, world, world))
        args.gpus = world
    if args.config == "C5":
        w = workloads.swarm_workload(args.particles or 100_000)
    else:
        w = workloads.config(args.config, num_samples=args.num_samples)
        if args.scaling == "weak" and args.gpus > 1 and w.d >= 2:
            # per-GPU work fixed: axis 1 (slowest in the reference row order) gets N times the points
            per_axis = w.samples_per_axis
            per_axis[1] *= args.gpus
            w.num_samples = per_axis
            w.name = "%s x%d (weak)" % (w.name, args.gpus)
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return 0
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rc = 1
    try:
        rc = run_b200(args, w, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            try:
                dist.barrier()
            finally:
                dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
