"""Where a sharded SafeOpt.optimize() step spends its time (run under torchrun, or alone): host-clock split of the step into
update_confidence_intervals / compute_sets / get_new_query_point, the K2 device time, and the in-kernel phase stamps of the
fused set pass on every rank.  Usage: torchrun ... tools/time_step_multi.py [--config C4] [--steps 20]"""
import argparse, os, sys, time
os.environ["SO_FUSED_DEBUG_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import safeopt_b200 as sb
from safeopt_b200 import workloads

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4")
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = workloads.config(args.config)
grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
       for i in range(w.n_gps)]
opt = sb.SafeOpt(gps if w.n_gps > 1 else gps[0], grid, w.fmin if w.n_gps > 1 else w.fmin[0], beta=w.beta, threshold=w.threshold)
for _ in range(5):
    opt.optimize()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
T = np.zeros(4)
k2 = []
stamps = []
for it in range(args.steps):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opt.update_confidence_intervals()
    e1.record()
    t1 = time.perf_counter()
    opt.compute_sets()
    t2 = time.perf_counter()
    opt.get_new_query_point()
    t3 = time.perf_counter()
    T += [t1 - t0, t2 - t1, t3 - t2, t3 - t0]
    k2.append(e0.elapsed_time(e1))
    st = np.zeros(16, dtype=np.int64)
    opt._engine.lib.so_debug_fused_times(opt._engine.handle, st.ctypes.data)
    stamps.append(st.copy())
T = T / args.steps * 1e6
st = np.array(stamps, dtype=float)
d = lambda a, b: float(np.mean(st[:, b] - st[:, a])) / 1e3
line = ("rank %d: step %.0f us = update_ci(host) %.0f + compute_sets %.0f + query %.0f | K2 device %.0f us | fused kernel us: scanA %.1f pubA %.1f waitA %.1f scanB %.1f "
        "pubB %.1f waitB %.1f scanC %.1f arrive %.1f pubwaitC %.1f copy %.1f total %.1f" % (
            rank, T[3], T[0], T[1], T[2], 1e3 * np.mean(k2), d(0, 1), d(1, 3), d(3, 4), d(4, 5), d(5, 7), d(7, 8), d(8, 9), d(9, 10), d(10, 11), d(11, 12), d(0, 12)))
for r in range(world):
    if r == rank:
        print(line, flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
