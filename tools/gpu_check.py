"""Developer diagnostic: run each stage of the CUDA path against the oracle and print error levels.
(Not a test -- tests/ holds the asserted versions.)  Usage: python tools/gpu_check.py [quick]"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import safeopt_b200 as sb  # noqa: E402
from safeopt_b200 import _lib, workloads  # noqa: E402
from safeopt_b200.engine import DeviceEngine  # noqa: E402
from oracle import gpy_lite as GPy, safeopt_port as port  # noqa: E402

KERN = {0: GPy.kern.RBF, 1: GPy.kern.Matern32, 2: GPy.kern.Matern52}
SBK = {0: sb.RBF, 1: sb.Matern32, 2: sb.Matern52}


def stage(name):
    def deco(fn):
        def run(*a, **k):
            t0 = time.time()
            try:
                fn(*a, **k)
                print("[ok  ] %-40s %.2fs" % (name, time.time() - t0), flush=True)
            except Exception:
                print("[FAIL] %s" % name, flush=True)
                traceback.print_exc()
        return run
    return deco


@stage("fit vs numpy")
def check_fit():
    for N, d, kind in [(5, 1, 0), (64, 2, 0), (100, 3, 1), (256, 4, 0), (300, 2, 2), (512, 6, 0)]:
        rs = np.random.RandomState(N)
        X = rs.uniform(-1.5, 1.5, (N, d))
        Y = rs.randn(N)
        ls = rs.uniform(0.7, 1.5, d)
        eng = DeviceEngine(max_gps=1)
        eng.fit(0, X, Y, kind, ls, 2.0, 0.05 ** 2)
        L, Linv, alpha = eng.fit_export(0, N)
        gp = GPy.models.GPRegression(X, Y[:, None], kernel=KERN[kind](d, variance=2.0, lengthscale=ls, ARD=True), noise_var=0.05 ** 2)
        Lr = gp.woodbury_chol
        print("   N=%d d=%d kind=%d  |L-Lref|=%.2e  |Linv L - I|=%.2e  |alpha-ref|/|alpha|=%.2e" % (
            N, d, kind, np.abs(L - np.tril(Lr)).max(), np.abs(Linv @ L - np.eye(N)).max(),
            np.abs(alpha - gp.woodbury_vector[:, 0]).max() / np.abs(alpha).max()))
        eng.close()


@stage("posterior rows (DMMA) vs simple vs oracle")
def check_posterior():
    for N, d, kind, M in [(5, 1, 0, 100), (64, 2, 0, 5000), (100, 3, 1, 3000), (128, 2, 2, 4097), (256, 4, 0, 20000),
                          (300, 2, 0, 3000), (512, 6, 0, 4000), (40, 2, 0, 70000)]:
        rs = np.random.RandomState(N + 1)
        X = rs.uniform(-1.5, 1.5, (N, d))
        Y = 2 * np.exp(-np.sum(X * X, 1) / 8) + 0.05 * rs.randn(N)
        ls = rs.uniform(0.7, 1.5, d)
        Xs = rs.uniform(-5, 5, (M, d))
        eng = DeviceEngine(max_gps=1)
        eng.fit(0, X, Y, kind, ls, 2.0, 0.05 ** 2)
        Xd = eng.to_device(Xs)
        mean, var = eng.empty((M,)), eng.empty((M,))
        Q = eng.empty((M, 2))
        S = eng.zeros((M,), "u8")
        eng.posterior_rows(0, Xd, 2.0, 0.0, mean=mean, var=var, Q=Q, q_col=0, S=S, safe_mode=_lib.SAFE_WRITE)
        ms, vs = eng.posterior_rows_simple(0, Xd)
        gp = GPy.models.GPRegression(X, Y[:, None], kernel=KERN[kind](d, variance=2.0, lengthscale=ls, ARD=True), noise_var=0.05 ** 2)
        mo, vo = gp.predict_noiseless(Xs)
        mean, var, ms, vs = [t.cpu().numpy() for t in (mean, var, ms, vs)]
        Qh = Q.cpu().numpy()
        Qo = np.stack([mo[:, 0] - 2 * np.sqrt(vo[:, 0]), mo[:, 0] + 2 * np.sqrt(vo[:, 0])], 1)
        print("   N=%d d=%d kind=%d M=%d  mean: dmma-simple %.1e dmma-oracle %.1e | var: dmma-simple %.1e dmma-oracle %.1e | Q %.1e | S mism %d (minmargin %.1e)" % (
            N, d, kind, M, np.abs(mean - ms).max(), np.abs(mean - mo[:, 0]).max(), np.abs(var - vs).max(),
            np.abs(var - vo[:, 0]).max(), np.abs(Qh - Qo).max(), int(((Qo[:, 0] > 0) != S.cpu().numpy().astype(bool)).sum()),
            np.abs(Qo[:, 0]).min()))
        eng.close()


@stage("posterior grid vs rows")
def check_grid():
    for d, n, N in [(1, 100, 5), (2, 60, 64), (3, [7, 9, 11], 30), (4, 12, 256)]:
        w = workloads.grid_workload("t", d, n if np.isscalar(n) else 10, N)
        bounds = w.bounds
        grid = sb.linearly_spaced_combinations(bounds, n)
        from safeopt_b200.utilities import detect_grid
        axes = detect_grid(grid)
        eng = DeviceEngine(max_gps=1)
        eng.fit(0, w.X, w.Y[:, 0], 0, w.lengthscale, w.variance, w.noise_var)
        eng.define_grid(axes)
        eng.prepare_grid(0)
        M = grid.shape[0]
        rows = eng.grid_rows(0, M).cpu().numpy()
        assert np.array_equal(rows, grid), "grid_rows mismatch"
        m1, v1, m2, v2 = eng.empty((M,)), eng.empty((M,)), eng.empty((M,)), eng.empty((M,))
        eng.posterior_grid(0, 0, M, 2.0, 0.0, mean=m1, var=v1)
        eng.posterior_rows(0, eng.to_device(grid), 2.0, 0.0, mean=m2, var=v2)
        print("   d=%d M=%d N=%d  grid-vs-rows mean %.1e var %.1e" % (d, M, N, (m1 - m2).abs().max().item(), (v1 - v2).abs().max().item()))
        # shard offset
        h = M // 3
        m3 = eng.empty((M - h,))
        eng.posterior_grid(0, h, M - h, 2.0, 0.0, mean=m3)
        assert torch.equal(m3, m1[h:]), "row0 offset mismatch"
        eng.close()


def make_pair(w, kind=0):
    gps_o = [GPy.models.GPRegression(w.X, w.Y[:, [i]], kernel=KERN[kind](w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var) for i in range(w.n_gps)]
    gps_d = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=SBK[kind](w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var) for i in range(w.n_gps)]
    return gps_o, gps_d


@stage("SafeOpt.optimize vs port (named configs, reduced grids)")
def check_optimize():
    for name, ns, explicit in [("C1", None, False), ("C2", 60, False), ("C2", 60, True), ("C3", 60, False), ("C4", 8, False), ("C2", 200, False)]:
        w = workloads.config(name, num_samples=ns)
        grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
        if explicit:
            grid = grid + 0.0
            grid[0, 0] += 1e-12
        go, gd = make_pair(w)
        p = port.GridProblem.create(go, grid, w.fmin, beta=w.beta, threshold=w.threshold)
        xo, rowo = p.optimize()
        opt = sb.SafeOpt(gd if len(gd) > 1 else gd[0], grid, w.fmin if len(gd) > 1 else w.fmin[0], beta=w.beta, threshold=w.threshold)
        xd = opt.optimize()
        print("   %s M=%d explicit=%s  grid_path=%s  |Q-Qo|=%.1e  S/M/G mism %d/%d/%d  row %d vs %d  nS=%d nM=%d nG=%d" % (
            name, grid.shape[0], explicit, opt._grid_axes is not None, np.abs(opt.Q - p.Q).max(), (opt.S != p.S).sum(), (opt.M != p.M).sum(),
            (opt.G != p.G).sum(), opt.last_query_row, rowo, p.S.sum(), p.M.sum(), p.G.sum()))
        mo = port.current_maximum(grid, p.Q, p.S)
        md = opt.get_maximum()
        print("      get_maximum: %s vs %s ; ucb row: %s" % (md[1], mo[1], np.array_equal(opt.optimize(ucb=True), port.new_query_point(grid, p.Q, p.S, p.M, p.G, p.scaling, ucb=True)[0])))


def expander_problem(seed=0, d=2, n=40, N=40, spread=2.5, fmin=0.5, thr=0.05, G=1):
    rs = np.random.RandomState(seed)
    X = rs.uniform(-spread, spread, size=(N, d))
    f = 2 * np.exp(-np.sum(X * X, 1) / 8)
    Y = np.stack([f + 0.05 * np.random.RandomState(seed + i + 1).randn(N) for i in range(G)], 1)
    return X, Y, sb.linearly_spaced_combinations([(-5, 5)] * d, n), [fmin] * G, thr


@stage("expander search vs port (refit-based)")
def check_expanders():
    for (spread, fmin, G, seed) in [(2.5, 0.5, 1, 0), (1.5, 1.5, 1, 0), (2.5, 0.5, 2, 0), (2.5, 0.8, 1, 3), (2.0, 1.0, 2, 5)]:
        X, Y, grid, fm, thr = expander_problem(seed=seed, spread=spread, fmin=fmin, G=G)
        d = X.shape[1]
        go = [GPy.models.GPRegression(X, Y[:, [i]], kernel=GPy.kern.RBF(d, variance=2.0, lengthscale=np.ones(d), ARD=True), noise_var=0.05 ** 2) for i in range(G)]
        gd = [sb.GPRegression(X, Y[:, [i]], kernel=sb.RBF(d, variance=2.0, lengthscale=np.ones(d), ARD=True), noise_var=0.05 ** 2) for i in range(G)]
        p = port.GridProblem.create(go, grid, fm, beta=2.0, threshold=thr)
        tr = {}
        xo, rowo = p.optimize(trace=tr)
        opt = sb.SafeOpt(gd, grid, fm, beta=2.0, threshold=thr)
        xd = opt.optimize()
        print("   spread=%.1f fmin=%.1f G=%d: cand %s/%s visited %s/%s  G rows %s vs %s  query row %d vs %d  order-eq %s" % (
            spread, fmin, G, opt.last_trace.get("n_candidates"), tr.get("n_candidates"), opt.last_trace.get("visited"), tr.get("visited"),
            np.flatnonzero(opt.G), np.flatnonzero(p.G), opt.last_query_row, rowo,
            None if "order" not in tr else np.array_equal(tr["order"][:tr["visited"]], opt.last_trace["order"][:tr["visited"]])))
        # full sets on a coarser grid
    X, Y, grid, fm, thr = expander_problem(spread=2.5, fmin=0.5, n=20)
    go = [GPy.models.GPRegression(X, Y[:, [0]], kernel=GPy.kern.RBF(2, variance=2.0, lengthscale=np.ones(2), ARD=True), noise_var=0.05 ** 2)]
    gd = [sb.GPRegression(X, Y[:, [0]], kernel=sb.RBF(2, variance=2.0, lengthscale=np.ones(2), ARD=True), noise_var=0.05 ** 2)]
    Qo = port.confidence_intervals(go, grid, 2.0)
    So, Mo, Go = port.compute_sets(go, grid, Qo, np.array(fm), 2.0, np.sqrt([2.0]), thr, full_sets=True)
    opt = sb.SafeOpt(gd, grid, fm, beta=2.0, threshold=thr)
    opt.update_confidence_intervals()
    opt.compute_sets(full_sets=True)
    print("   full_sets: G mism %d of %d (nG=%d)" % ((opt.G != Go).sum(), Go.size, Go.sum()))


@stage("BO loop trajectory vs port")
def check_loop():
    X, Y, grid, fm, thr = expander_problem(seed=1, spread=1.0, fmin=0.3, N=6, n=30)
    f = lambda x: 2 * np.exp(-np.sum(np.atleast_2d(x) ** 2, 1) / 8)
    go = [GPy.models.GPRegression(X, Y[:, [0]], kernel=GPy.kern.RBF(2, variance=2.0, lengthscale=np.ones(2), ARD=True), noise_var=0.05 ** 2)]
    gd = [sb.GPRegression(X, Y[:, [0]], kernel=sb.RBF(2, variance=2.0, lengthscale=np.ones(2), ARD=True), noise_var=0.05 ** 2)]
    p = port.GridProblem.create(go, grid, fm, beta=2.0, threshold=thr)
    opt = sb.SafeOpt(gd, grid, fm, beta=2.0, threshold=thr)
    same = 0
    for it in range(25):
        xo, ro = p.optimize()
        xd = opt.optimize()
        if ro != opt.last_query_row:
            print("   diverged at iteration %d: %d vs %d (nG %d vs %d)" % (it, ro, opt.last_query_row, p.G.sum(), opt.G.sum()))
            break
        same += 1
        y = f(xo)[:, None]
        p.add_new_data_point(xo, y)
        opt.add_new_data_point(xd, y)
    print("   %d identical iterations; final N=%d; expander iterations seen: yes" % (same, opt.t))


@stage("swarm fitness vs port")
def check_swarm():
    w = workloads.swarm_workload(n_particles=5000, n_train=60)
    go = [GPy.models.GPRegression(w.X, w.Y[:, [i]], kernel=GPy.kern.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var) for i in range(w.n_gps)]
    gd = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var) for i in range(w.n_gps)]
    fmin = [0.0, 0.2]
    opt = sb.SafeOptSwarm(gd, fmin, bounds=w.bounds, beta=2.0, swarm_size=20)
    sc = opt.scaling
    for kind in ["greedy", "maximizers", "expanders", "safe_set"]:
        vo, so = port.particle_fitness(go, np.array(fmin), 2.0, sc, kind, w.particles, best_lower_bound=0.7)
        opt.best_lower_bound = 0.7
        vd, sd = opt._compute_particle_fitness(kind, w.particles)
        print("   %-10s |v-vo| %.2e (scale %.2e)  safe mism %d" % (kind, np.abs(vd - vo).max(), np.abs(vo).max(), (sd != so).sum()))
    np.random.seed(0)
    x = opt.optimize()
    print("   SafeOptSwarm.optimize ->", x, "S size", opt.S.shape)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    check_fit()
    check_posterior()
    check_grid()
    check_optimize()
    check_expanders()
    check_loop()
    check_swarm()
