"""Generate tests/golden/*.npz by running the UNMODIFIED reference package (imported from
/root/reference through oracle.compat) over oracle.gpy_lite.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the reference tree does not exist on the GPU box):

    OPENBLAS_NUM_THREADS=1 python -m oracle.make_golden

Every fixture stores its inputs (training data, hyper-parameters, grid description, thresholds)
next to the reference's outputs (Q, masks, query row/point, maximum), so the GPU tests can rebuild
the problem without the reference.  Masks are stored as packed bits.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

KINDS = {"rbf": 0, "mat32": 1, "mat52": 2}


def kernel_of(GPy, kind, d, variance, ls):
    cls = {"rbf": GPy.kern.RBF, "mat32": GPy.kern.Matern32, "mat52": GPy.kern.Matern52}[kind]
    return cls(d, variance=variance, lengthscale=np.asarray(ls, dtype=float), ARD=True)


def row_of(grid, x):
    hit = np.flatnonzero(np.all(grid == x[None, :], axis=1))
    return int(hit[0])


def grid_case(ref, GPy, name, X, Y, kind, variance, ls, noise, bounds, n, fmin, beta, threshold, full_sets=False, lipschitz=None):
    d = X.shape[1]
    G = Y.shape[1]
    gps = [GPy.models.GPRegression(X, Y[:, [i]], kernel=kernel_of(GPy, kind, d, variance, ls), noise_var=noise) for i in range(G)]
    grid = ref.linearly_spaced_combinations(bounds, n)
    opt = ref.SafeOpt(gps if G > 1 else gps[0], grid, fmin=list(fmin) if G > 1 else fmin[0], beta=beta, threshold=threshold,
                      lipschitz=lipschitz)
    if full_sets:
        opt.update_confidence_intervals()
        opt.compute_sets(full_sets=True)
        x_next = opt.get_new_query_point()
    else:
        x_next = opt.optimize()
    S, M, Gm = opt.S.copy(), opt.M.copy(), opt.G.copy()
    Q = opt.Q.copy()
    row = row_of(grid, np.atleast_1d(x_next))
    mx = opt.get_maximum()
    x_ucb = opt.optimize(ucb=True)
    # smallest distances of any bound to a decision threshold (hard part 2)
    margin_S = float(np.min(np.abs(Q[:, ::2] - np.asarray(fmin)[None, :])))
    margin_M = float(np.min(np.abs(Q[S, 1] - np.max(Q[S, 0])))) if S.any() else np.inf
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), X=X, Y=Y, kind=KINDS[kind], variance=variance, lengthscale=np.asarray(ls, dtype=float),
        noise_var=noise, bounds=np.asarray(bounds, dtype=float), num_samples=np.asarray(n if not np.isscalar(n) else [n] * d),
        fmin=np.asarray(fmin, dtype=float), beta=beta, threshold=threshold, full_sets=full_sets,
        lipschitz=np.asarray([] if lipschitz is None else np.atleast_1d(lipschitz), dtype=float),
        Q=Q, S=np.packbits(S), M=np.packbits(M), G=np.packbits(Gm), n_rows=grid.shape[0],
        x_next=np.atleast_1d(x_next), row_next=row, max_x=np.atleast_1d(mx[0]), max_val=float(mx[1]),
        row_ucb=row_of(grid, np.atleast_1d(x_ucb)), margin_S=margin_S, margin_M=margin_M)
    print("%-28s M=%-7d N=%-4d G=%d  |S|=%d |M|=%d |G|=%d row=%d  margins S %.1e M %.1e" % (
        name, grid.shape[0], X.shape[0], G, S.sum(), M.sum(), Gm.sum(), row, margin_S, margin_M))


def synth(seed, N, d, G, spread, noise_sd=0.05):
    rs = np.random.RandomState(seed)
    X = rs.uniform(-spread, spread, size=(N, d))
    f = 2.0 * np.exp(-np.sum(X * X, axis=1) / 8.0)
    Y = np.stack([f + noise_sd * np.random.RandomState(seed + 1 + i).randn(N) for i in range(G)], axis=1)
    return X, Y


def loop_case(ref, GPy, name, iters=20):
    """A short Bayesian-optimisation run: the sequence of query rows is the fixture."""
    X, Y = synth(1, 6, 2, 1, 1.0)
    bounds, n = [(-5.0, 5.0)] * 2, 30
    gp = GPy.models.GPRegression(X, Y, kernel=kernel_of(GPy, "rbf", 2, 2.0, [1.0, 1.0]), noise_var=0.05 ** 2)
    grid = ref.linearly_spaced_combinations(bounds, n)
    opt = ref.SafeOpt(gp, grid, fmin=0.3, beta=2.0, threshold=0.05)
    rows, ys, ng = [], [], []
    for _ in range(iters):
        x = opt.optimize()
        rows.append(row_of(grid, x))
        ng.append(int(opt.G.sum()))
        y = 2.0 * np.exp(-np.sum(x * x) / 8.0)
        ys.append(y)
        opt.add_new_data_point(x, np.array([[y]]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, Y=Y, bounds=np.asarray(bounds), num_samples=n, fmin=0.3, beta=2.0,
                        threshold=0.05, variance=2.0, lengthscale=np.ones(2), noise_var=0.05 ** 2, rows=np.asarray(rows),
                        ys=np.asarray(ys), n_expanders=np.asarray(ng))
    print("%-28s rows %s  expander iterations %d" % (name, rows[:8], int(np.sum(np.asarray(ng) > 0))))


def swarm_case(ref, GPy, name):
    rs = np.random.RandomState(7)
    d, N, G, P = 3, 40, 2, 600
    X = rs.uniform(-0.5, 0.5, size=(N, d))
    f = 1.0 - 0.3 * np.sum(X * X, axis=1)
    Y = np.stack([f + 0.05 * np.random.RandomState(8 + i).randn(N) for i in range(G)], axis=1)
    gps = [GPy.models.GPRegression(X, Y[:, [i]], kernel=kernel_of(GPy, "rbf", d, 2.0, np.ones(d)), noise_var=0.05 ** 2) for i in range(G)]
    fmin = [0.0, 0.2]
    opt = ref.SafeOptSwarm(gps, fmin, bounds=[(-1.0, 1.0)] * d, beta=2.0, swarm_size=20)
    opt.best_lower_bound = 0.7
    particles = rs.uniform(-1.0, 1.0, size=(P, d))
    out = {}
    for kind in ["greedy", "maximizers", "expanders", "safe_set"]:
        v, s = opt._compute_particle_fitness(kind, particles)
        out["values_" + kind] = np.asarray(v, dtype=float)
        out["safe_" + kind] = np.asarray(np.broadcast_to(s, (P,)), dtype=bool)
    pen_in = np.array([0.5, 0.0, -0.0005, -0.001, -0.05, -0.1, -0.5, -1.0, -1.5, -3.0])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, Y=Y, fmin=np.asarray(fmin), particles=particles, beta=2.0,
                        variance=2.0, lengthscale=np.ones(d), noise_var=0.05 ** 2, best_lower_bound=0.7,
                        velocities=opt.optimal_velocities, scaling=opt.scaling, penalty_in=pen_in,
                        penalty_out=opt._compute_penalty(pen_in.copy()), **out)
    print("%-28s velocities %s" % (name, opt.optimal_velocities))


def swarm_query_case(ref, GPy, name, kind="rbf", swarm_size=48, iters=25):
    """SafeOptSwarm.optimize() unrolled (gp_opt.py:1136-1177) with the global NumPy stream seeded: per swarm
    the safe set before/after, the swarm's best positions (input of the insertion step, gp_opt.py:1088-1110),
    the returned point and value.  The GPU tests replay the insertion on the stored best positions and the
    whole trajectory with the same seed."""
    rs = np.random.RandomState(11)
    d, N, G = 2, 12, 2
    X = rs.uniform(-0.4, 0.4, size=(N, d))
    f = 1.0 - 0.3 * np.sum(X * X, axis=1)
    Y = np.stack([f + 0.05 * np.random.RandomState(12 + i).randn(N) for i in range(G)], axis=1)
    ls = np.array([0.8, 1.1])
    gps = [GPy.models.GPRegression(X, Y[:, [i]], kernel=kernel_of(GPy, kind, d, 2.0, ls), noise_var=0.05 ** 2) for i in range(G)]
    fmin = [0.0, 0.2]
    opt = ref.SafeOptSwarm(gps, fmin, bounds=[(-1.5, 1.5)] * d, beta=2.0, swarm_size=swarm_size)
    opt.max_iters = iters
    np.random.seed(5)
    out = {}
    for stage in ["greedy", "maximizers", "expanders"]:
        out[stage + "_S_before"] = opt.S.copy()
        x, v = opt.get_new_query_point(stage)
        if stage == "greedy":
            opt.greedy, opt.best_lower_bound = x, v
        out[stage + "_x"] = np.asarray(x, dtype=float)
        out[stage + "_v"] = np.asarray(v, dtype=float)
        out[stage + "_best_positions"] = opt.swarms[stage].best_positions.copy()
        out[stage + "_best_values"] = opt.swarms[stage].best_values.copy()
        out[stage + "_S_after"] = opt.S.copy()
        # the insertion must not have been preceded by a pruning of S in this fixture
        assert np.array_equal(out[stage + "_S_after"][:out[stage + "_S_before"].shape[0]], out[stage + "_S_before"])
    # decision margins of the insertion step (knife-edge fixtures would be useless for parity)
    from oracle import safeopt_port as port
    margins = []
    for stage in ["maximizers", "expanders"]:
        acc, margin = port.select_new_safe_points(gps[0].kern, out[stage + "_S_before"], out[stage + "_best_positions"], opt.scaling[0])
        assert np.array_equal(out[stage + "_best_positions"][acc], out[stage + "_S_after"][out[stage + "_S_before"].shape[0]:])
        margins.append(margin)
    # a full optimize() from the same seed (fresh optimiser)
    gps2 = [GPy.models.GPRegression(X, Y[:, [i]], kernel=kernel_of(GPy, kind, d, 2.0, ls), noise_var=0.05 ** 2) for i in range(G)]
    opt2 = ref.SafeOptSwarm(gps2, fmin, bounds=[(-1.5, 1.5)] * d, beta=2.0, swarm_size=swarm_size)
    opt2.max_iters = iters
    np.random.seed(5)
    x_next = opt2.optimize()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, Y=Y, fmin=np.asarray(fmin), beta=2.0, variance=2.0, lengthscale=ls,
                        noise_var=0.05 ** 2, kind=KINDS[kind], bounds=np.asarray([(-1.5, 1.5)] * d), swarm_size=swarm_size,
                        max_iters=iters, seed=5, scaling=opt.scaling, insertion_margin=np.asarray(margins), x_next=np.asarray(x_next),
                        S_final=opt2.S, **out)
    print("%-28s |S| %d -> %d -> %d, insertion margins %s, next %s" % (
        name, out["maximizers_S_before"].shape[0], out["maximizers_S_after"].shape[0], out["expanders_S_after"].shape[0], margins, x_next))


def context_case(ref, GPy, name, lipschitz=None, threshold=5.0):
    """examples/context_example.ipynb shape: 1 parameter + 1 context, product of RBF kernels on disjoint dims.
    (With the GP-based expander test and candidates present the reference raises IndexError -- gp_opt.py:585-588
    hands a 1-D x to _add_context; fixtures therefore use the Lipschitz rule or a threshold that leaves no candidates.)"""
    rs = np.random.RandomState(21)
    N = 12
    X = np.hstack([rs.uniform(-1.0, 1.0, (N, 1)), rs.uniform(0.0, 1.0, (N, 1))])
    f = np.exp(-X[:, 0] ** 2) * (1.0 + 0.5 * X[:, 1])
    Y = (f + 0.02 * rs.randn(N))[:, None]
    kern = GPy.kern.RBF(1, variance=2.0, lengthscale=0.7, active_dims=[0]) * GPy.kern.RBF(1, variance=1.5, lengthscale=2.0, active_dims=[1])
    gp = GPy.models.GPRegression(X, Y, kernel=kern, noise_var=0.02 ** 2)
    pset = ref.linearly_spaced_combinations([(-2.0, 2.0)], 300)
    opt = ref.SafeOpt(gp, pset, fmin=0.2, num_contexts=1, beta=2.0, threshold=threshold, lipschitz=lipschitz)
    out = {}
    for k, ctx in enumerate([0.25, 0.9]):
        x = opt.optimize(context=np.array([ctx]))
        out["ctx%d" % k] = ctx
        out["Q%d" % k] = opt.Q.copy()
        out["S%d" % k], out["M%d" % k], out["G%d" % k] = np.packbits(opt.S), np.packbits(opt.M), np.packbits(opt.G)
        out["x%d" % k] = np.atleast_1d(x)
        mx = opt.get_maximum(context=np.array([ctx]))
        out["maxx%d" % k], out["maxv%d" % k] = np.atleast_1d(mx[0]), float(mx[1])
        print("%-28s ctx=%.2f |S|=%d |M|=%d |G|=%d x=%s" % (name, ctx, opt.S.sum(), opt.M.sum(), opt.G.sum(), x))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, Y=Y, pset=pset, fmin=0.2, beta=2.0, threshold=threshold,
                        lipschitz=np.asarray([] if lipschitz is None else [lipschitz], dtype=float),
                        var0=2.0, ls0=0.7, var1=1.5, ls1=2.0, noise_var=0.02 ** 2, n_rows=pset.shape[0], **out)


def main():
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    from oracle import compat, gpy_lite as GPy
    ref = compat.import_reference()
    sys.path.insert(0, os.path.join(ROOT))
    from safeopt_b200 import workloads

    # closed-form doctest configuration (gp_opt.py:327-338)
    grid_case(ref, GPy, "doctest_1d", np.array([[0.0]]), np.array([[1.0]]), "rbf", 1.0, [1.0], 0.01 ** 2, [(-1.0, 1.0)], 100, [0.0], 2.0, 0)
    # named configurations (reduced grids where the full one would be too large for a fixture)
    for name, ns in [("C1", None), ("C2", None), ("C3", 80), ("C4", 10)]:
        w = workloads.config(name, num_samples=ns)
        grid_case(ref, GPy, "config_%s" % name + ("" if ns is None else "_n%d" % ns), w.X, w.Y, "rbf", w.variance, w.lengthscale, w.noise_var,
                  w.bounds, w.num_samples, w.fmin, w.beta, w.threshold)
    # expander-exercising problems
    X, Y = synth(0, 40, 2, 1, 2.5)
    grid_case(ref, GPy, "expander_g1", X, Y, "rbf", 2.0, [1.0, 1.0], 0.05 ** 2, [(-5.0, 5.0)] * 2, 40, [0.5], 2.0, 0.05)
    X, Y = synth(0, 40, 2, 2, 2.5)
    grid_case(ref, GPy, "expander_g2", X, Y, "rbf", 2.0, [1.0, 1.0], 0.05 ** 2, [(-5.0, 5.0)] * 2, 40, [0.5, 0.5], 2.0, 0.05)
    X, Y = synth(0, 40, 2, 1, 1.5)
    grid_case(ref, GPy, "expander_tight", X, Y, "rbf", 2.0, [1.0, 1.0], 0.05 ** 2, [(-5.0, 5.0)] * 2, 40, [1.5], 2.0, 0.05)
    X, Y = synth(0, 40, 2, 1, 2.5)
    grid_case(ref, GPy, "full_sets_g1", X, Y, "rbf", 2.0, [1.0, 1.0], 0.05 ** 2, [(-5.0, 5.0)] * 2, 20, [0.5], 2.0, 0.05, full_sets=True)
    # Matern kernels, ARD lengthscales, ragged axes, unconstrained objective (fmin = -inf for GP 0)
    X, Y = synth(3, 50, 3, 1, 2.0)
    grid_case(ref, GPy, "matern32_3d", X, Y, "mat32", 1.5, [0.8, 1.3, 1.0], 0.02, [(-3.0, 3.0), (-2.0, 2.0), (-1.0, 4.0)], [9, 14, 11], [0.2], 3.0, 0.1)
    X, Y = synth(4, 33, 2, 2, 2.0)
    grid_case(ref, GPy, "matern52_2d_g2", X, Y, "mat52", 2.0, [1.2, 0.7], 0.05 ** 2, [(-4.0, 4.0)] * 2, [31, 45], [-np.inf, 0.4], 2.0, 0.05)
    loop_case(ref, GPy, "bo_loop_2d")
    swarm_case(ref, GPy, "swarm_fitness_3d")
    swarm_query_case(ref, GPy, "swarm_query_2d")
    swarm_query_case(ref, GPy, "swarm_query_2d_mat32", kind="mat32", swarm_size=40, iters=15)
    # Lipschitz expander rule (gp_opt.py:558-576) and contexts (gp_opt.py:424-451)
    X, Y = synth(0, 40, 2, 1, 2.5)
    grid_case(ref, GPy, "lipschitz_g1", X, Y, "rbf", 2.0, [1.0, 1.0], 0.05 ** 2, [(-5.0, 5.0)] * 2, 40, [0.5], 2.0, 0.05, lipschitz=2.0)
    X, Y = synth(0, 40, 2, 2, 2.5)
    grid_case(ref, GPy, "lipschitz_g2", X, Y, "rbf", 2.0, [1.0, 1.0], 0.05 ** 2, [(-5.0, 5.0)] * 2, 40, [0.5, 0.5], 2.0, 0.05,
              lipschitz=[5.0, 0.1])
    context_case(ref, GPy, "context_1p1c")
    context_case(ref, GPy, "context_1p1c_lipschitz", lipschitz=1.5, threshold=0.05)


if __name__ == "__main__":
    main()
