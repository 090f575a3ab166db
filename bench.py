#!/usr/bin/env python
"""Benchmark of the SafeOpt hot path (BASELINE.json metric: grid-point GP posterior + safe-set evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one ``SafeOpt.optimize()`` over the whole candidate grid: GP posterior of every row
(kernel rows, L^-1 contraction, mean/var), confidence bounds, safe set, maximisers, expander
candidates / search and the query-point argmax.  Workload = BASELINE config 4 (the configuration
the metric is quoted on): d=4, 50^4 = 6.25e6 rows, N_train=256, 1 GP (objective = constraint),
fp64, synthetic RBF problem of SURVEY.md section 8d.  With N ranks the candidate rows are split into
N contiguous row blocks.  ``--scaling weak`` (default): every rank keeps a full config-4 block, i.e.
the grid becomes 50 x (50 N) x 50 x 50 (axis 1 is the slowest axis of the reference's row order, so
rank r owns axis-1 indices [50 r, 50 (r+1))) -- at N=1 this IS config 4.  ``--scaling strong``: the
fixed 50^4 grid is split N ways (the config's "sharded 8xB200" reading).

Reported: ``value`` (device-resident throughput, fit cached), ``e2e`` (through the public API
with host inputs: refit from host X/Y every step + optimize + result read-back), ``roofline`` of
the dominant kernel (fp64 tensor pipe; peak = cuBLAS DGEMM measured in this session, because
MEASURED_PEAKS.json only carries HBM and bf16 numbers), ``cpu_baseline`` (oracle port on the
host cores on a bounded sample of the same workload) and the clocks seen during timing.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4"])
    ap.add_argument("--num-samples", type=int, default=None, help="override points per axis (development only)")
    ap.add_argument("--cpu-sample-rows", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--explicit-rows", action="store_true", help="force the explicit-rows kernel path (no grid tables)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: per-GPU rows fixed (grid axis 1 grows with N); strong: the 50^4 grid is split N ways")
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
    """Time the oracle port (NumPy/SciPy restatement of the reference path) on a bounded row sample."""
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
    return take, times


def run_reference(args, w, rank, world):
    """--impl reference: the reference algorithm's CPU path (oracle port) on the host cores."""
    if rank != 0:
        return
    cores = os.cpu_count()
    take, times = cpu_port_run(w, args.cpu_sample_rows, steps=args.steps, warmup=args.warmup)
    total = float(np.sum(times))
    value = take * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, args, grid_path=None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d contiguous grid rows of the %d-row workload per step (oracle/safeopt_port.py over "
                                   "oracle/gpy_lite.py, NumPy/OpenBLAS threads=%d, 100k-row chunks)" % (take, w.n_rows, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(w, args, grid_path):
    return {"workload": "%s: %dD RBF-ARD, %d constraint GP(s), %s=%d grid rows, N_train=%d, %s, beta=%g, fmin=0, threshold=%g" % (
        w.name, w.d, w.n_gps, "x".join(str(n) for n in w.samples_per_axis), w.n_rows, w.n_train, w.dtype, w.beta, w.threshold),
        "rows": w.n_rows, "rows_per_gpu": -(-w.n_rows // args.gpus), "n_train": w.n_train, "d": w.d, "n_gps": w.n_gps,
        "parallelism": "rows sharded x%d (%s scaling)" % (args.gpus, args.scaling),
        "candidate_path": None if grid_path is None else ("grid rows generated on device" if grid_path else "explicit rows in HBM"),
        "l2": "each step streams 33 B/row of outputs (206 MB at C4, > 126 MB L2) and re-derives everything else; "
              "operands (L^-1, tables, 0.8 MB) are meant to stay cache-resident"}


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args, w, rank, world, local_rank):
    import torch
    import safeopt_b200 as sb
    from safeopt_b200 import _lib
    from safeopt_b200.gp_opt import _DeviceFits  # noqa: F401

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod

    grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
    if args.explicit_rows:
        os.environ["SAFEOPT_B200_GRID_FAST_PATH"] = "0"      # explicit-rows kernels (200 MB of candidates in HBM)
    gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                           noise_var=w.noise_var, device=dev) for i in range(w.n_gps)]
    opt = sb.SafeOpt(gps if w.n_gps > 1 else gps[0], grid, w.fmin if w.n_gps > 1 else w.fmin[0], beta=w.beta,
                     threshold=w.threshold, device=dev)
    grid_path = opt._grid_axes is not None
    eng = opt._engine

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: fit cached, everything else per step
    for _ in range(args.warmup):
        x_next = opt.optimize()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    k2_events = []
    orig_pg, orig_pr, orig_pm = eng.posterior_grid, eng.posterior_rows, eng.posterior_multi

    def timed(fn):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            k2_events.append((e0, e1))
            return out
        return wrapper

    eng.posterior_grid, eng.posterior_rows, eng.posterior_multi = timed(orig_pg), timed(orig_pr), timed(orig_pm)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        x_next = opt.optimize()
    ev1.record()
    barrier()
    eng.posterior_grid, eng.posterior_rows, eng.posterior_multi = orig_pg, orig_pr, orig_pm
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launches - launches0
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in k2_events]))
    clocks = sampler.stop() if rank == 0 else None
    n_safe, n_max = opt._safe_info["n_safe"], opt._max_info["n_max"] if opt._max_info else 0
    trace = dict(opt.last_trace)

    # ---- end to end through the public API with host inputs: refit from host X/Y + optimize + result to host
    X_pin = torch.from_numpy(np.ascontiguousarray(w.X)).pin_memory()
    Y_pin = torch.from_numpy(np.ascontiguousarray(w.Y)).pin_memory()
    h2d = X_pin.numel() * 8 + Y_pin.numel() * 8 + (w.d + 2) * 8 * w.n_gps
    d2h = 136 * world + w.n_gps * 4    # every rank's two 64-byte records + candidate count (one copy), fit status words

    def e2e_step():
        Xh, Yh = X_pin.numpy(), Y_pin.numpy()
        for i, gp in enumerate(opt.gps):
            gp.set_XY(Xh, Yh[:, [i]])
        opt._fits.invalidate()          # a new observation would change the fingerprint; force the refit
        return np.asarray(opt.optimize())

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        x_e2e = e2e_step()
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0)))

    rows = w.n_rows
    value = rows * args.steps / (dev_ms * 1e-3)
    e2e_value = rows * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (k_posterior): fp64 tensor pipe
    local_rows = opt._row1 - opt._row0
    fpe = flops_per_eval(w.n_train, w.d, 1)
    achieved = fpe * local_rows / (k2_ms * 1e-3) / 1e12
    peak, peak_src = measure_dgemm_tflops(torch, dev) if rank == 0 else (None, None)
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = None
    if os.path.exists(peaks_file):
        try:
            hbm_peak = json.load(open(peaks_file)).get("hbm_gbs")
        except Exception:
            pass
    if rank != 0:
        return
    traffic = None
    tf = os.path.join(ROOT, "profiles", "k_posterior_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "tensor", "kernel": "k_posterior (DMMA.8x8x4 fp64 tensor pipe)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                "peak_source": peak_src, "flops_per_eval": fpe, "evals_per_launch": local_rows,
                "kernel_ms_per_launch": k2_ms, "kernel_launches_per_step": len(k2_events) / args.steps,
                "kernel_share_of_step": k2_ms * (len(k2_events) / args.steps) / (dev_ms / args.steps),
                "fp64_dmma_microbench_tflops": 37.1, "algorithmic_hbm_gbs": bytes_per_eval(w.d, w.n_gps, grid_path) * local_rows / (k2_ms * 1e-3) / 1e9,
                "hbm_peak_gbs_measured": hbm_peak,
                "traffic_note": "684 MB of the measured DRAM traffic is the precomputed scaled-operand table A'(s) (2500 x 270 KB), "
                                "read exactly once per launch; it replaces 1.6e9 fp64 exp per launch on the FP64 pipe the "
                                "contraction saturates. Outputs: 33 B/row = 206 MB",
                "note": "fp64 path: MEASURED_PEAKS.json has no fp64 number, so the denominator is cuBLAS DGEMM measured in this session"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        take, times = cpu_port_run(w, args.cpu_sample_rows, steps=1, warmup=0)
        cpu = {"value": take / times[0], "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "%d contiguous grid rows of the %d-row workload, one optimize() (oracle/safeopt_port.py over "
                         "oracle/gpy_lite.py, NumPy/OpenBLAS default threads, 100k-row chunks), %.1f s" % (take, rows, times[0])}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(w, args, grid_path),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms / args.steps,
                "what": "set_XY from pinned host X/Y -> device refit (Cholesky, L^-1, tables) -> optimize() -> next parameters on host"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "result": {"x_next": [float(v) for v in np.asarray(x_next)], "n_safe": int(n_safe), "n_maximizers": int(n_max),
                   "n_expander_candidates": int(trace.get("n_candidates", 0)), "e2e_same_x": bool(np.array_equal(x_next, x_e2e))},
    }
    print(json.dumps(line), flush=True)


def measure_dgemm_tflops(torch, dev, n=8192):
    """cuBLAS DGEMM rate in this session = the fp64 roofline denominator (burst, timed alone)."""
    try:
        a = torch.zeros((n, n), dtype=torch.float64, device=dev)
        b = torch.zeros((n, n), dtype=torch.float64, device=dev)
        torch.matmul(a, b)
        torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        del a, b
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "torch.matmul fp64 %d^3 (cuBLAS DGEMM), best of 3, this session" % n
    except Exception as exc:  # pragma: no cover
        return 35.76, "fallback: profiles/r01_fp64_rates_b200.jsonl cublas_dgemm n=8192 (%s)" % exc


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
    w = workloads.config(args.config, num_samples=args.num_samples)
    if args.scaling == "weak" and args.gpus > 1 and w.d >= 2:
        # per-GPU work fixed: axis 1 (slowest in the reference row order) gets N times the points
        per_axis = w.samples_per_axis
        per_axis[1] *= args.gpus
        w.num_samples = per_axis
        w.name = "%s x%d (weak)" % (w.name, args.gpus)
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, w, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
