"""CPU tests of the product's host orchestration (safeopt_b200/gp_opt.py) over a NumPy stand-in for the device engine
(tests/fake_engine.py, built on the oracle): the class surface must reproduce the reference's golden results when the
device calls honour the C-ABI contracts -- fit bookkeeping, chained set passes, record parsing, candidate ordering, the
batched expander search, G bookkeeping, query-point selection and the BO loop."""
import numpy as np
import pytest

from conftest import GRID_CASES, golden_lipschitz, golden_problem, load_golden, unpack_mask
from fake_engine import FakeEngine

import safeopt_b200 as sb
from safeopt_b200 import gp_opt


@pytest.fixture()
def fake_device(monkeypatch):
    monkeypatch.setattr(gp_opt, "DeviceEngine", FakeEngine)
    monkeypatch.delenv("SAFEOPT_B200_GRID_FAST_PATH", raising=False)
    monkeypatch.delenv("SAFEOPT_B200_SHARE_FITS", raising=False)
    return FakeEngine


@pytest.mark.parametrize("name", [c for c in GRID_CASES if c not in ("config_C2",)])      # C2: 40k rows x N = 64, slow in NumPy
def test_class_surface_reproduces_golden(name, fake_device):
    g = load_golden(name)
    gps, grid, fmin = golden_problem(g, "gpu")
    n_rows = int(g["n_rows"])
    opt = sb.SafeOpt(gps if len(gps) > 1 else gps[0], grid, fmin if len(gps) > 1 else fmin[0], lipschitz=golden_lipschitz(g),
                     beta=float(g["beta"]), threshold=float(g["threshold"]))
    full = bool(g["full_sets"]) if "full_sets" in g.files else False
    if full:
        opt.update_confidence_intervals()
        opt.compute_sets(full_sets=True)
    else:
        x = opt.optimize()
        assert opt.last_query_row == int(g["row_next"]) and np.array_equal(x, g["x_next"])
    assert np.abs(opt.Q - g["Q"]).max() < 1e-9
    assert np.array_equal(opt.S, unpack_mask(g["S"], n_rows))
    assert np.array_equal(opt.M, unpack_mask(g["M"], n_rows))
    assert np.array_equal(opt.G, unpack_mask(g["G"], n_rows))
    mx = opt.get_maximum()
    assert np.array_equal(mx[0], g["max_x"]) and abs(mx[1] - float(g["max_val"])) < 1e-9
    assert opt.optimize(ucb=True) is not None and opt.last_query_row == int(g["row_ucb"])


def test_bo_loop_with_incremental_fits(fake_device):
    g = load_golden("bo_loop_2d")
    gp = sb.GPRegression(g["X"], g["Y"], kernel=sb.RBF(2, variance=2.0, lengthscale=np.ones(2), ARD=True), noise_var=float(g["noise_var"]))
    grid = sb.linearly_spaced_combinations([tuple(b) for b in g["bounds"]], int(g["num_samples"]))
    opt = sb.SafeOpt(gp, grid, float(g["fmin"]), beta=float(g["beta"]), threshold=float(g["threshold"]))
    for it, row_ref in enumerate(g["rows"]):
        x = opt.optimize()
        assert opt.last_query_row == int(row_ref), "trajectory diverged at iteration %d" % it
        assert int(opt.G.sum()) == int(g["n_expanders"][it])
        opt.add_new_data_point(x, np.array([[g["ys"][it]]]))
    kinds = [c[0] for c in opt._engine.calls if c[0] in ("fit", "append", "remove")]
    assert kinds == ["fit"] + ["append"] * (len(g["rows"]) - 1)
    # every (re)fit is followed by the grid tables of that GP for this rank's rows
    assert sum(c[0] == "prepare" for c in opt._engine.calls) == len(g["rows"])


def test_groups_share_one_launch(fake_device):
    g = load_golden("config_C3_n80")
    gps, grid, fmin = golden_problem(g, "gpu")
    opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
    opt.optimize()
    assert [c for c in opt._engine.calls if c[0] == "multi"] == [("multi", (0, 1, 2))]
    assert opt.last_query_row == int(g["row_next"])


def test_no_safe_points(fake_device):
    gp = sb.GPRegression(np.array([[0.0]]), np.array([[-1.0]]), noise_var=0.01 ** 2)
    opt = sb.SafeOpt(gp, sb.linearly_spaced_combinations([(-1.0, 1.0)], 50), 0.0)
    with pytest.raises(EnvironmentError):
        opt.optimize()
    assert opt.get_maximum() is None and not opt.S.any() and not opt.M.any() and not opt.G.any()


@pytest.mark.parametrize("name", ["context_1p1c", "context_1p1c_lipschitz"])
def test_contexts(name, fake_device):
    """Contexts (gp_opt.py:424-451): product of RBF kernels merged into one ARD kernel by extract_hyper, explicit rows,
    context column rewritten on the device copy."""
    g = load_golden(name)
    kern = sb.RBF(1, variance=float(g["var0"]), lengthscale=float(g["ls0"]), active_dims=[0]) * \
        sb.RBF(1, variance=float(g["var1"]), lengthscale=float(g["ls1"]), active_dims=[1])
    gp = sb.GPRegression(g["X"], g["Y"], kernel=kern, noise_var=float(g["noise_var"]))
    lip = g["lipschitz"]
    opt = sb.SafeOpt(gp, g["pset"], float(g["fmin"]), num_contexts=1, beta=float(g["beta"]), threshold=float(g["threshold"]),
                     lipschitz=None if lip.size == 0 else float(lip[0]))
    n_rows = int(g["n_rows"])
    with pytest.raises(ValueError):
        opt.optimize()
    for k in range(2):
        ctx = np.array([float(g["ctx%d" % k])])
        x = opt.optimize(context=ctx)
        assert np.abs(opt.Q - g["Q%d" % k]).max() < 1e-9
        assert np.array_equal(opt.S, unpack_mask(g["S%d" % k], n_rows))
        assert np.array_equal(opt.M, unpack_mask(g["M%d" % k], n_rows))
        assert np.array_equal(opt.G, unpack_mask(g["G%d" % k], n_rows))
        assert np.array_equal(x, g["x%d" % k]) and x.shape == (1,)
    opt.add_new_data_point(x, np.array([[0.5]]), context=ctx)
    assert opt.x.shape == (13, 2) and opt.x[-1, 1] == ctx[0]


@pytest.mark.parametrize("name", ["swarm_query_2d", "swarm_query_2d_mat32"])
def test_safeoptswarm_trajectory(name, fake_device):
    """SafeOptSwarm with the host swarm back end over the stand-in engine follows the reference's seeded trajectory:
    safe-set re-check, particle sampling, PSO, correlation-filtered growth of the safe set, greedy-point bookkeeping."""
    g = load_golden(name)
    X, Y = g["X"], g["Y"]
    d = X.shape[1]
    cls = {0: sb.RBF, 1: sb.Matern32, 2: sb.Matern52}[int(g["kind"])]
    gps = [sb.GPRegression(X, Y[:, [i]], kernel=cls(d, variance=float(g["variance"]), lengthscale=g["lengthscale"], ARD=True),
                           noise_var=float(g["noise_var"])) for i in range(Y.shape[1])]
    opt = sb.SafeOptSwarm(gps, list(g["fmin"]), bounds=[tuple(b) for b in g["bounds"]], beta=float(g["beta"]),
                          swarm_size=int(g["swarm_size"]), swarm_backend="host")
    opt.max_iters = int(g["max_iters"])
    assert np.allclose(opt.optimal_velocities, np.asarray(opt.optimal_velocities)) and opt.swarm_backend == "host"
    np.random.seed(int(g["seed"]))
    for stage in ["greedy", "maximizers", "expanders"]:
        assert np.array_equal(opt.S, g[stage + "_S_before"]), stage
        x, v = opt.get_new_query_point(stage)
        if stage == "greedy":
            opt.greedy, opt.best_lower_bound = x, v
        assert np.abs(opt.swarms[stage].best_positions - g[stage + "_best_positions"]).max() < 1e-9, stage
        assert np.abs(np.asarray(x) - g[stage + "_x"]).max() < 1e-9 and np.abs(np.asarray(v) - g[stage + "_v"]).max() < 1e-9, stage
        assert opt.S.shape == g[stage + "_S_after"].shape and np.abs(opt.S - g[stage + "_S_after"]).max() < 1e-9, stage
