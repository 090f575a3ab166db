"""CPU tests that pin the oracle: closed forms, an independent second implementation (scikit-learn),
the reference's own package (build container only) and the committed golden fixtures."""
import numpy as np
import pytest

from conftest import GRID_CASES, golden_lipschitz, golden_problem, load_golden, oracle_kernel, unpack_mask
from oracle import gpy_lite, safeopt_port as port


def test_closed_form_single_point():
    """N=1 doctest configuration (gp_opt.py:327-338): mu = k/(1+s2+1e-8), var = 1 - k^2/(1+s2+1e-8)."""
    gp = gpy_lite.GPRegression(np.array([[0.0]]), np.array([[1.0]]), noise_var=0.01 ** 2)
    x = np.linspace(-1, 1, 100)[:, None]
    k = np.exp(-x[:, 0] ** 2 / 2)
    den = 1 + 1e-4 + 1e-8
    mean, var = gp.predict_noiseless(x)
    assert np.abs(mean[:, 0] - k / den).max() < 1e-14
    assert np.abs(var[:, 0] - (1 - k * k / den)).max() < 1e-14
    assert mean.shape == (100, 1) and var.shape == (100, 1)


def test_closed_form_matern_two_points():
    """N=2 Matern32 / Matern52 by hand-solved 2x2 system."""
    X = np.array([[0.0], [1.0]])
    Y = np.array([[1.0], [0.5]])
    xs = np.array([[0.3], [2.0]])
    for cls, prof in [(gpy_lite.Matern32, lambda r: (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r)),
                      (gpy_lite.Matern52, lambda r: (1 + np.sqrt(5) * r + 5 / 3 * r * r) * np.exp(-np.sqrt(5) * r))]:
        v, ell, s2 = 1.7, 0.8, 0.01
        gp = gpy_lite.GPRegression(X, Y, kernel=cls(1, variance=v, lengthscale=ell), noise_var=s2)
        kf = lambda a, b: v * prof(np.abs(a - b) / ell)
        K = np.array([[kf(0, 0) + s2 + 1e-8, kf(0, 1)], [kf(1, 0), kf(1, 1) + s2 + 1e-8]])
        det = K[0, 0] * K[1, 1] - K[0, 1] * K[1, 0]
        Ki = np.array([[K[1, 1], -K[0, 1]], [-K[1, 0], K[0, 0]]]) / det
        mean, var = gp.predict_noiseless(xs)
        for j, x in enumerate(xs[:, 0]):
            kx = np.array([kf(x, 0.0), kf(x, 1.0)])
            assert abs(mean[j, 0] - kx @ Ki @ Y[:, 0]) < 1e-13
            assert abs(var[j, 0] - (v - kx @ Ki @ kx)) < 1e-13


def test_kernel_surface():
    """Kdiag = variance (pinned by the reference's own test_gps.py:48-60), K symmetric with that diagonal."""
    X = np.random.RandomState(0).randn(7, 3)
    for cls in (gpy_lite.RBF, gpy_lite.Matern32, gpy_lite.Matern52):
        k = cls(3, variance=2.5, lengthscale=[0.5, 1.0, 2.0], ARD=True)
        K = k.K(X)
        assert np.allclose(np.diag(K), 2.5) and np.allclose(K, K.T)
        assert np.allclose(k.Kdiag(X), 2.5)
        assert np.allclose(k.K(X, X[:2]), K[:, :2], atol=1e-14)
    prod = gpy_lite.RBF(1, variance=2.0, active_dims=[0]) * gpy_lite.RBF(1, variance=3.0, lengthscale=0.5, active_dims=[1])
    X2 = X[:, :2]
    merged = gpy_lite.RBF(2, variance=6.0, lengthscale=[1.0, 0.5], ARD=True)
    assert np.allclose(prod.K(X2), merged.K(X2), atol=1e-13)


def test_against_sklearn():
    """Independent implementation: scikit-learn's GaussianProcessRegressor with alpha = noise + 1e-8."""
    sk = pytest.importorskip("sklearn.gaussian_process")
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    rs = np.random.RandomState(5)
    X = rs.uniform(-1.5, 1.5, (64, 2))
    Y = (2 * np.exp(-np.sum(X * X, 1) / 8) + 0.05 * rs.randn(64))[:, None]
    Xs = rs.uniform(-5, 5, (3600, 2))
    ls = np.array([0.9, 1.3])
    for ours, theirs in [(gpy_lite.RBF(2, variance=2.0, lengthscale=ls, ARD=True), ConstantKernel(2.0) * RBF(ls)),
                         (gpy_lite.Matern32(2, variance=2.0, lengthscale=ls, ARD=True), ConstantKernel(2.0) * Matern(ls, nu=1.5)),
                         (gpy_lite.Matern52(2, variance=2.0, lengthscale=ls, ARD=True), ConstantKernel(2.0) * Matern(ls, nu=2.5))]:
        gp = gpy_lite.GPRegression(X, Y, kernel=ours, noise_var=0.05 ** 2)
        reg = sk.GaussianProcessRegressor(kernel=theirs, alpha=0.05 ** 2 + 1e-8, optimizer=None).fit(X, Y[:, 0])
        m2, s2 = reg.predict(Xs, return_std=True)
        m1, v1 = gp.predict_noiseless(Xs)
        assert np.abs(m1[:, 0] - m2).max() < 1e-10
        assert np.abs(v1[:, 0] - s2 ** 2).max() < 1e-9 * 2.0


def test_grid_order_matches_reference_description():
    """Row order of linearly_spaced_combinations: axis 1 slowest, axis 0 next, axes 2.. fastest."""
    g = port.linearly_spaced_combinations([(0, 1), (10, 12), (100, 103)], [2, 3, 4])
    assert g.shape == (24, 3)
    assert np.array_equal(g[:4, 2], [100, 101, 102, 103]) and np.all(g[:4, 0] == 0) and np.all(g[:8, 1] == 10)
    assert g[4, 0] == 1 and g[8, 1] == 11
    g1 = port.linearly_spaced_combinations([(-1, 1)], 5)
    assert g1.shape == (5, 1) and g1[-1, 0] == 1.0


def test_grid_rows_equal_full_grid():
    """port.grid_rows (used by the CPU baseline to take a sample without building a 50M-row grid) is bit-identical to
    the rows of linearly_spaced_combinations."""
    for bounds, n in [([(-5, 5)] * 4, [5, 10, 3, 4]), ([(-1, 2)], 7), ([(-5, 5), (0, 1)], [4, 6]), ([(-2, 2)] * 3, 5)]:
        full = port.linearly_spaced_combinations(bounds, n)
        rows = np.arange(full.shape[0])[::-1]
        assert np.array_equal(port.grid_rows(bounds, n, rows), full[rows])


def test_penalty_matches_reference_table():
    g = load_golden("swarm_fitness_3d")
    assert np.array_equal(port.penalty(g["penalty_in"]), g["penalty_out"])


@pytest.mark.parametrize("name", GRID_CASES)
def test_port_reproduces_golden(name):
    """The port, on the fixture's inputs, reproduces what the reference produced (bit-identical Q)."""
    g = load_golden(name)
    n_rows = int(g["n_rows"])
    if n_rows > 20000:
        pytest.skip("large fixture is covered by the GPU parity test")
    gps, grid, fmin = golden_problem(g, "cpu")
    beta, thr = float(g["beta"]), float(g["threshold"])
    prob = port.GridProblem.create(gps, grid, fmin if len(fmin) > 1 else fmin[0], beta=beta, threshold=thr,
                                   lipschitz=golden_lipschitz(g))
    if bool(g["full_sets"]):
        prob.Q = port.confidence_intervals(prob.gps, grid, beta)
        prob.S, prob.M, prob.G = port.compute_sets(prob.gps, grid, prob.Q, prob.fmin, beta, prob.scaling, thr, full_sets=True)
        x, row = port.new_query_point(grid, prob.Q, prob.S, prob.M, prob.G, prob.scaling)
    else:
        x, row = prob.optimize()
    assert np.array_equal(prob.Q, g["Q"])
    assert np.array_equal(prob.S, unpack_mask(g["S"], n_rows))
    assert np.array_equal(prob.M, unpack_mask(g["M"], n_rows))
    assert np.array_equal(prob.G, unpack_mask(g["G"], n_rows))
    assert row == int(g["row_next"]) and np.array_equal(x, g["x_next"])
    mx = port.current_maximum(grid, prob.Q, prob.S)
    assert np.array_equal(mx[0], g["max_x"]) and mx[1] == float(g["max_val"])
    _, row_ucb = port.new_query_point(grid, prob.Q, prob.S, prob.M, prob.G, prob.scaling, ucb=True)
    assert row_ucb == int(g["row_ucb"])


def test_port_swarm_fitness_golden():
    g = load_golden("swarm_fitness_3d")
    X, Y = g["X"], g["Y"]
    d = X.shape[1]
    gps = [gpy_lite.GPRegression(X, Y[:, [i]], kernel=gpy_lite.RBF(d, variance=2.0, lengthscale=np.ones(d), ARD=True),
                                 noise_var=float(g["noise_var"])) for i in range(Y.shape[1])]
    for kind in ["greedy", "maximizers", "expanders", "safe_set"]:
        v, s = port.particle_fitness(gps, g["fmin"], float(g["beta"]), g["scaling"], kind, g["particles"],
                                     best_lower_bound=float(g["best_lower_bound"]))
        assert np.allclose(v, g["values_" + kind], rtol=1e-13, atol=1e-13)
        assert np.array_equal(s, g["safe_" + kind])


@pytest.mark.parametrize("name", ["swarm_query_2d", "swarm_query_2d_mat32"])
def test_port_swarm_insertion_golden(name):
    """select_new_safe_points == the reference's safe-set growth (gp_opt.py:1088-1110) on the stored swarms."""
    g = load_golden(name)
    kern = oracle_kernel(int(g["kind"]), 2, float(g["variance"]), g["lengthscale"])
    for stage in ["maximizers", "expanders"]:
        before, after, best = g[stage + "_S_before"], g[stage + "_S_after"], g[stage + "_best_positions"]
        acc, margin = port.select_new_safe_points(kern, before, best, float(g["scaling"][0]))
        assert np.array_equal(best[acc], after[before.shape[0]:])
        assert margin > 1e-6


def test_port_bo_loop_golden():
    g = load_golden("bo_loop_2d")
    gp = gpy_lite.GPRegression(g["X"], g["Y"], kernel=gpy_lite.RBF(2, variance=2.0, lengthscale=np.ones(2), ARD=True),
                               noise_var=float(g["noise_var"]))
    grid = port.linearly_spaced_combinations([tuple(b) for b in g["bounds"]], int(g["num_samples"]))
    prob = port.GridProblem.create([gp], grid, float(g["fmin"]), beta=float(g["beta"]), threshold=float(g["threshold"]))
    for it, row_ref in enumerate(g["rows"]):
        x, row = prob.optimize()
        assert row == int(row_ref), "diverged at iteration %d" % it
        assert int(prob.G.sum()) == int(g["n_expanders"][it])
        prob.add_new_data_point(x, np.array([[g["ys"][it]]]))


# ---------------------------------------------------------------- against the reference itself (build container only)
def test_port_matches_reference_live(reference_pkg):
    ref = reference_pkg
    rs = np.random.RandomState(11)
    for G, fmin, n in [(1, 0.4, 25), (2, 0.5, 30)]:
        X = rs.uniform(-2.5, 2.5, (35, 2))
        f = 2 * np.exp(-np.sum(X * X, 1) / 8)
        Y = np.stack([f + 0.05 * rs.randn(35) for _ in range(G)], 1)
        mk = lambda: [gpy_lite.GPRegression(X, Y[:, [i]], kernel=gpy_lite.RBF(2, variance=2.0, lengthscale=[1.0, 1.2], ARD=True),
                                            noise_var=0.05 ** 2) for i in range(G)]
        grid = ref.linearly_spaced_combinations([(-5, 5)] * 2, n)
        assert np.array_equal(grid, port.linearly_spaced_combinations([(-5, 5)] * 2, n))
        a = mk()
        opt = ref.SafeOpt(a if G > 1 else a[0], grid, fmin=[fmin] * G if G > 1 else fmin, beta=2.0, threshold=0.05)
        x_ref = opt.optimize()
        prob = port.GridProblem.create(mk(), grid, [fmin] * G if G > 1 else fmin, beta=2.0, threshold=0.05)
        x, row = prob.optimize()
        assert np.array_equal(opt.Q, prob.Q)
        assert np.array_equal(opt.S, prob.S) and np.array_equal(opt.M, prob.M) and np.array_equal(opt.G, prob.G)
        assert np.array_equal(x_ref, x)


def test_reference_own_tests_surface(reference_pkg):
    """The reference's unit-test expectations (safeopt/tests/test_gps.py:48-60) hold over gpy_lite."""
    ref = reference_pkg
    from safeopt.gp_opt import GaussianProcessOptimization
    gp1 = gpy_lite.GPRegression(np.array([[0.0]]), np.array([[0.0]]), kernel=gpy_lite.RBF(1, variance=2))
    gp2 = gpy_lite.GPRegression(np.array([[0.0]]), np.array([[0.0]]), kernel=gpy_lite.Matern32(1, variance=4))
    opt = GaussianProcessOptimization([gp1, gp2], fmin=0, beta=2, num_contexts=1, threshold=0, scaling="auto")
    assert np.allclose(opt.scaling, [np.sqrt(2), np.sqrt(4)])
    assert ref.SafeOpt is not None
