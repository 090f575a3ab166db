"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs, against the committed golden fixtures produced by the
reference itself, and -- at BASELINE.json's full sizes -- through size-independent properties.

Tolerances (fp64; SURVEY.md section 7, hard part 1): |d mean| <= 1e-11 * max|y|, |d var| <= 1e-10 * k(x,x),
|d l|, |d u| <= 1e-9 * sqrt(k(x,x)); masks and query rows bit-exact (fixture margins are >= 1e-5)."""
import numpy as np
import pytest

from conftest import GRID_CASES, device_kernel, golden_lipschitz, golden_problem, load_golden, oracle_kernel, unpack_mask

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover - the CPU run deselects this module with -m "not gpu"
    pytest.skip("no CUDA device", allow_module_level=True)

import safeopt_b200 as sb  # noqa: E402
from safeopt_b200 import _lib, workloads  # noqa: E402
from safeopt_b200 import distributed as D  # noqa: E402
from safeopt_b200.engine import DeviceEngine  # noqa: E402
from oracle import gpy_lite, safeopt_port as port  # noqa: E402


def _problem(N, d, seed):
    rs = np.random.RandomState(seed)
    X = rs.uniform(-1.5, 1.5, (N, d))
    Y = 2 * np.exp(-np.sum(X * X, 1) / 8) + 0.05 * rs.randn(N)
    ls = rs.uniform(0.7, 1.5, d)
    return X, Y, ls, rs


# ---------------------------------------------------------------- K1
@pytest.mark.parametrize("N,d,kind", [(1, 1, 0), (5, 1, 0), (7, 2, 1), (33, 2, 0), (64, 2, 0), (100, 3, 1), (256, 4, 0), (300, 2, 2), (512, 6, 0),
                                      (1000, 3, 0), (2048, 4, 0)])
def test_fit_matches_lapack(N, d, kind):
    X, Y, ls, _ = _problem(N, d, N)
    eng = DeviceEngine(max_gps=1)
    eng.fit(0, X, Y, kind, ls, 2.0, 0.05 ** 2)
    L, Linv, alpha = eng.fit_export(0, N)
    gp = gpy_lite.GPRegression(X, Y[:, None], kernel=oracle_kernel(kind, d, 2.0, ls), noise_var=0.05 ** 2)
    assert np.abs(L - np.tril(gp.woodbury_chol)).max() < 1e-12
    assert np.abs(Linv @ L - np.eye(N)).max() < 1e-11
    assert np.abs(alpha - gp.woodbury_vector[:, 0]).max() < 1e-10 * max(1.0, np.abs(alpha).max())
    eng.close()


@pytest.mark.parametrize("N,d,kind", [(1, 1, 0), (31, 2, 1), (32, 2, 0), (33, 3, 0), (200, 2, 2), (256, 4, 0), (257, 4, 0), (511, 3, 0), (512, 6, 1)])
def test_one_launch_cluster_fit_equals_kernel_per_panel_fit(N, d, kind, monkeypatch):
    """k_fit_cluster (one launch, DMMA tiles, cluster barriers) against the kernel-per-panel factorisation and LAPACK."""
    X, Y, ls, _ = _problem(N, d, 3 * N + d)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SO_FIT_CLUSTER", "16" if mode == "1" else "0")
        eng = DeviceEngine(max_gps=1)
        eng.fit(0, X, Y, kind, ls, 2.0, 0.05 ** 2)
        res[mode] = eng.fit_export(0, N)
        Xq = np.random.RandomState(5).uniform(-2, 2, (200, d))
        mean, var = eng.empty((200,)), eng.empty((200,))
        eng.posterior_rows(0, eng.to_device(Xq), 2.0, -np.inf, mean=mean, var=var)
        res[mode] += (mean.cpu().numpy(), var.cpu().numpy())
        eng.close()
    gp = gpy_lite.GPRegression(X, Y[:, None], kernel=oracle_kernel(kind, d, 2.0, ls), noise_var=0.05 ** 2)
    L, Linv, alpha, mean, var = res["1"]
    assert np.abs(L - np.tril(gp.woodbury_chol)).max() < 1e-12
    assert np.abs(Linv @ L - np.eye(N)).max() < 1e-11
    assert np.abs(alpha - gp.woodbury_vector[:, 0]).max() < 1e-10 * max(1.0, np.abs(alpha).max())
    assert np.abs(L - res["0"][0]).max() < 1e-12 and np.abs(Linv - res["0"][1]).max() < 1e-9 * np.abs(Linv).max()
    assert np.abs(mean - res["0"][3]).max() < 1e-10 and np.abs(var - res["0"][4]).max() < 1e-10


@pytest.mark.parametrize("N0,steps,d,kind", [(1, 20, 1, 0), (5, 30, 2, 0), (60, 12, 3, 1), (255, 4, 4, 0), (500, 3, 6, 2)])
def test_fit_append_and_remove_match_refit(N0, steps, d, kind):
    """f4: one-point appends (bordered Cholesky, gp_opt.py:227) and removals (:267, :275) reproduce the
    from-scratch factorisation of the same data; buffers are grown by a refit when the capacity is hit."""
    import scipy.linalg as sla
    X, Y, ls, _ = _problem(N0 + steps, d, seed=7 + N0)
    var, noise = 2.0, 0.05 ** 2
    eng = DeviceEngine(max_gps=1)
    eng.fit(0, X[:N0], Y[:N0], kind, ls, var, noise)
    kern = oracle_kernel(kind, d, var, ls)
    Xq = np.random.RandomState(1).uniform(-2, 2, size=(300, d))
    n = N0
    appended = 0
    for _ in range(steps):
        if eng.fit_append(0, X[n], Y[n]):
            appended += 1
        else:                                   # capacity: the host's fallback
            eng.fit(0, X[:n + 1], Y[:n + 1], kind, ls, var, noise)
        n += 1
        if n in (N0 + 1, N0 + steps) or n % 8 in (0, 1):
            L, Linv, alpha = eng.fit_export(0, n)
            Ky = kern.K(X[:n]) + (noise + 1e-8) * np.eye(n)
            Lr = np.linalg.cholesky(Ky)
            assert np.abs(L - Lr).max() < 1e-11 * np.abs(Lr).max()
            Lir = sla.solve_triangular(Lr, np.eye(n), lower=True)
            assert np.abs(Linv - Lir).max() < 1e-9 * np.abs(Lir).max()
            assert np.abs(alpha - sla.cho_solve((Lr, True), Y[:n])).max() < 1e-8 * max(1.0, np.abs(alpha).max())
            go = gpy_lite.GPRegression(X[:n], Y[:n, None], kernel=kern, noise_var=noise)
            mo, vo = go.predict_noiseless(Xq)
            mean, varr = eng.empty((300,)), eng.empty((300,))
            eng.posterior_rows(0, eng.to_device(Xq), 2.0, -np.inf, mean=mean, var=varr)
            assert np.abs(mean.cpu().numpy() - mo[:, 0]).max() < 1e-9 * max(1.0, np.abs(Y).max())
            assert np.abs(varr.cpu().numpy() - vo[:, 0]).max() < 1e-9 * var
    assert appended >= steps - 2
    for _ in range(min(steps, 10)):
        eng.fit_remove_last(0)
        n -= 1
    go = gpy_lite.GPRegression(X[:n], Y[:n, None], kernel=kern, noise_var=noise)
    mo, vo = go.predict_noiseless(Xq)
    mean, varr = eng.empty((300,)), eng.empty((300,))
    eng.posterior_rows(0, eng.to_device(Xq), 2.0, -np.inf, mean=mean, var=varr)
    assert np.abs(mean.cpu().numpy() - mo[:, 0]).max() < 1e-9 * max(1.0, np.abs(Y).max())
    assert np.abs(varr.cpu().numpy() - vo[:, 0]).max() < 1e-9 * var
    _, _, alpha = eng.fit_export(0, n)
    assert np.abs(alpha - go.woodbury_vector[:, 0]).max() < 1e-8 * max(1.0, np.abs(alpha).max())
    eng.close()


def test_device_fits_use_incremental_updates():
    """add_new_data_point / remove_last_data_point reach the device as append / removal, anything else refits."""
    g = load_golden("expander_g1")
    gps, grid, fmin = golden_problem(g, "gpu")
    opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
    x = opt.optimize()
    assert (opt._fits.refits, opt._fits.appends) == (1, 0)
    opt.add_new_data_point(x, np.array([[1.0]]))
    x2 = opt.optimize()
    assert (opt._fits.refits, opt._fits.appends) == (1, 1)
    Q_inc = opt.Q.copy()
    opt._fits.invalidate()
    opt.optimize()
    assert opt._fits.refits == 2 and np.abs(opt.Q - Q_inc).max() < 1e-10
    opt.remove_last_data_point()
    assert np.array_equal(opt.optimize(), x) and opt._fits.removals == 1
    opt.gps[0].kern.variance = 2.5                   # hyper-parameter change: full refit
    opt.optimize()
    assert opt._fits.refits == 3


def test_fit_reports_not_positive_definite():
    eng = DeviceEngine(max_gps=1)
    X = np.zeros((3, 1))            # three identical points, zero noise -> singular even with the 1e-8 jitter in fp64? no: jitter keeps it PD
    eng.fit(0, X, np.zeros(3), 0, [1.0], 1.0, 0.0)
    Xbad = np.array([[0.0], [np.nan]])
    with pytest.raises(sb.DeviceError) as err:
        eng.fit(0, Xbad, np.zeros(2), 0, [1.0], 1.0, 0.0)
    assert err.value.status == _lib.SO_ERR_NOT_PD
    with pytest.raises(sb.DeviceError):
        eng.posterior_rows(0, eng.to_device(np.zeros((4, 1))), 2.0, 0.0, mean=eng.empty((4,)), var=eng.empty((4,)))
    with pytest.raises(sb.DeviceError):
        eng.fit(0, X, np.zeros(3), 7, [1.0], 1.0, 0.0)
    eng.close()


def test_failed_factorisation_surfaces_inside_optimize():
    """SafeOpt starts its fits without waiting (so_fit_async); a Cholesky that fails is still reported by the same optimize()
    call, when the records of the set pass arrive (and by the Q / S / M accessors)."""
    X = np.array([[0.0], [np.nan], [1.0]])
    gp = sb.GPRegression(X, np.zeros((3, 1)), kernel=sb.RBF(1, variance=1.0, lengthscale=1.0), noise_var=0.01)
    grid = sb.linearly_spaced_combinations([(-1, 1)], 50)
    opt = sb.SafeOpt(gp, grid, fmin=-1.0)
    assert opt._fits.async_fit
    with pytest.raises(sb.DeviceError) as err:
        opt.optimize()
    assert err.value.status == _lib.SO_ERR_NOT_PD
    opt2 = sb.SafeOpt(sb.GPRegression(X, np.zeros((3, 1)), kernel=sb.RBF(1), noise_var=0.01), grid, fmin=-1.0)
    opt2.update_confidence_intervals()
    with pytest.raises(sb.DeviceError):
        opt2.Q
    gp.set_XY(np.array([[0.0], [0.5], [1.0]]), np.ones((3, 1)))          # repaired data: the next call refits and works
    assert opt.optimize().shape == (1,)


# ---------------------------------------------------------------- K2
@pytest.mark.parametrize("N,d,kind,M", [(1, 1, 0, 100), (5, 1, 0, 100), (9, 2, 2, 1), (64, 2, 0, 5000), (100, 3, 1, 3000),
                                        (128, 2, 2, 4097), (256, 4, 0, 20000), (300, 2, 0, 3000), (512, 6, 0, 4000),
                                        (40, 2, 0, 70000), (700, 3, 1, 900), (1100, 3, 0, 600),
                                        (512, 6, 0, 20000), (300, 2, 1, 25000), (400, 3, 2, 30000)])      # generation split over all eight warps
def test_posterior_matches_oracle(N, d, kind, M):
    X, Y, ls, rs = _problem(N, d, N + 1)
    Xs = rs.uniform(-5, 5, (M, d))
    eng = DeviceEngine(max_gps=1)
    eng.fit(0, X, Y, kind, ls, 2.0, 0.05 ** 2)
    Xd = eng.to_device(Xs)
    mean, var, Q, S = eng.empty((M,)), eng.empty((M,)), eng.empty((M, 2)), eng.zeros((M,), "u8")
    eng.posterior_rows(0, Xd, 2.0, 0.0, mean=mean, var=var, Q=Q, q_col=0, S=S, safe_mode=_lib.SAFE_WRITE)
    ms, vs = eng.posterior_rows_simple(0, Xd)
    gp = gpy_lite.GPRegression(X, Y[:, None], kernel=oracle_kernel(kind, d, 2.0, ls), noise_var=0.05 ** 2)
    mo, vo = gp.predict_noiseless(Xs)
    mean, var, ms, vs, Q, S = [t.cpu().numpy() for t in (mean, var, ms, vs, Q, S)]
    scale_y = max(1.0, np.abs(Y).max())
    assert np.abs(mean - ms).max() < 1e-12 * scale_y and np.abs(var - vs).max() < 1e-12 * 2.0     # tensor-core vs plain DFMA kernel
    assert np.abs(mean - mo[:, 0]).max() < 1e-11 * scale_y
    assert np.abs(var - vo[:, 0]).max() < 1e-10 * 2.0
    Qo = np.stack([mo[:, 0] - 2 * np.sqrt(vo[:, 0]), mo[:, 0] + 2 * np.sqrt(vo[:, 0])], 1)
    assert np.abs(Q - Qo).max() < 1e-9 * np.sqrt(2.0) * 2
    # the device's own Q and S are exactly consistent; against the oracle only outside the tolerance band
    assert np.array_equal(S.astype(bool), Q[:, 0] > 0.0)
    clear = np.abs(Qo[:, 0]) > 1e-8
    assert np.array_equal(S.astype(bool)[clear], (Qo[:, 0] > 0.0)[clear])
    # report the margin (a seeded problem drifting onto a knife edge would otherwise pass silently through the band):
    # at most a handful of rows may sit inside the band, and the closest one is printed with -s
    print("posterior N=%d d=%d kind=%d M=%d: min |l - fmin| = %.3e, rows inside the 1e-8 band: %d" % (
        N, d, kind, M, np.abs(Qo[:, 0]).min(), int((~clear).sum())))
    assert (~clear).sum() <= max(1, M // 10000)
    assert np.array_equal(Q[:, 0], mean - 2.0 * np.sqrt(var)) and np.array_equal(Q[:, 1], mean + 2.0 * np.sqrt(var))
    eng.close()


def test_posterior_edge_shapes():
    """Largest input dimension (16), no candidates, one candidate, and the largest fit (N = 2048) on the explicit-rows path."""
    rs = np.random.RandomState(5)
    d, N = 16, 33
    X, Y, ls = rs.uniform(-1, 1, (N, d)), rs.randn(N), rs.uniform(1.0, 3.0, d)
    eng = DeviceEngine(max_gps=1)
    eng.fit(0, X, Y, 1, ls, 1.3, 0.01)
    gp = gpy_lite.GPRegression(X, Y[:, None], kernel=oracle_kernel(1, d, 1.3, ls), noise_var=0.01)
    for M in (0, 1, 777):
        Xs = rs.uniform(-1, 1, (M, d))
        mean, var = eng.empty((M,)), eng.empty((M,))
        before = eng.launches
        eng.posterior_rows(0, eng.to_device(Xs) if M else eng.empty((0, d)), 2.0, 0.0, mean=mean, var=var)
        if M == 0:
            assert eng.launches == before
            continue
        mo, vo = gp.predict_noiseless(Xs)
        assert np.abs(mean.cpu().numpy() - mo[:, 0]).max() < 1e-11 * max(1.0, np.abs(Y).max())
        assert np.abs(var.cpu().numpy() - vo[:, 0]).max() < 1e-10 * 1.3
    N = 2048
    X, Y, ls, _ = _problem(N, 4, 99)
    eng.fit(0, X, Y, 0, ls, 2.0, 0.05 ** 2)
    Xs = rs.uniform(-2, 2, (500, 4))
    mean, var = eng.empty((500,)), eng.empty((500,))
    try:
        eng.posterior_rows(0, eng.to_device(Xs), 2.0, 0.0, mean=mean, var=var)
    except sb.DeviceError as exc:               # the shared-memory tile caps the tensor-core kernel at N ~ 1600: refused, not wrong
        assert exc.status == _lib.SO_ERR_CAPACITY
    else:
        gp = gpy_lite.GPRegression(X, Y[:, None], kernel=oracle_kernel(0, 4, 2.0, ls), noise_var=0.05 ** 2)
        mo, vo = gp.predict_noiseless(Xs)
        assert np.abs(mean.cpu().numpy() - mo[:, 0]).max() < 1e-9 * max(1.0, np.abs(Y).max())
        assert np.abs(var.cpu().numpy() - vo[:, 0]).max() < 1e-9 * 2.0
    eng.close()


def test_safeopt_options_match_port():
    """Callable beta, per-GP threshold array and an explicit scaling list (gp_opt.py:74-99, :536) against the port
    (a plain list as threshold raises TypeError in the reference's `threshold * beta`; an array is what works there)."""
    g = load_golden("expander_g2")
    gps, grid, fmin = golden_problem(g, "gpu")
    gos, _, _ = golden_problem(g, "cpu")
    beta = lambda t: 1.5 + 0.01 * t
    kw = dict(beta=beta, threshold=np.array([0.05, 0.2]), scaling=[1.2, 0.9])
    opt = sb.SafeOpt(gps, grid, fmin, **kw)
    ref = port.GridProblem.create(gos, grid, fmin, **kw)
    x = opt.optimize()
    x_ref, row_ref = ref.optimize()
    assert np.abs(opt.Q - ref.Q).max() < 1e-9 * 2 * np.sqrt(2.0)
    assert np.array_equal(opt.S, ref.S) and np.array_equal(opt.M, ref.M) and np.array_equal(opt.G, ref.G)
    assert opt.last_query_row == row_ref and np.array_equal(x, x_ref)
    with pytest.raises(ValueError):
        sb.SafeOpt(gps, grid, fmin, scaling=[1.0])


def test_posterior_safe_modes_and_columns():
    X, Y, ls, rs = _problem(30, 2, 3)
    Xs = rs.uniform(-4, 4, (999, 2))
    eng = DeviceEngine(max_gps=2)
    eng.fit(0, X, Y, 0, ls, 2.0, 0.01)
    eng.fit(1, X, Y - 0.5, 1, ls, 1.0, 0.01)
    Xd = eng.to_device(Xs)
    Q = eng.zeros((999, 4))
    S = eng.zeros((999,), "u8")
    eng.posterior_rows(0, Xd, 2.0, 0.3, Q=Q, q_col=0, S=S, safe_mode=_lib.SAFE_WRITE)
    S0 = S.clone()
    eng.posterior_rows(1, Xd, 2.0, 0.1, Q=Q, q_col=2, S=S, safe_mode=_lib.SAFE_AND)
    Qh = Q.cpu().numpy()
    assert np.array_equal(S0.cpu().numpy().astype(bool), Qh[:, 0] > 0.3)
    assert np.array_equal(S.cpu().numpy().astype(bool), (Qh[:, 0] > 0.3) & (Qh[:, 2] > 0.1))
    eng.posterior_rows(1, Xd, 2.0, 0.1, Q=Q, q_col=2, S=S, safe_mode=_lib.SAFE_NONE)
    assert np.array_equal(S.cpu().numpy().astype(bool), (Qh[:, 0] > 0.3) & (Qh[:, 2] > 0.1))
    eng.posterior_rows(0, Xd[:0], 2.0, 0.0, Q=Q[:0], S=S[:0], safe_mode=_lib.SAFE_WRITE)      # empty input is a no-op
    eng.close()


@pytest.mark.parametrize("d,n,N", [(1, 100, 5), (2, 60, 64), (3, [7, 9, 11], 30), (4, 12, 256), (5, 4, 20), (6, 3, 40), (4, 12, 257), (2, 70, 300),
                                   (3, 20, 390), (2, 64, 520), (4, 12, 281), (4, 12, 330), (3, 40, 384)])
def test_grid_path_equals_rows_path(d, n, N):
    from safeopt_b200.utilities import detect_grid
    w = workloads.grid_workload("t", d, 10, N)
    grid = sb.linearly_spaced_combinations(w.bounds, n)
    axes = detect_grid(grid)
    eng = DeviceEngine(max_gps=1)
    eng.fit(0, w.X, w.Y[:, 0], 0, w.lengthscale, w.variance, w.noise_var)
    eng.define_grid(axes)
    eng.prepare_grid(0)
    M = grid.shape[0]
    assert np.array_equal(eng.grid_rows(0, M).cpu().numpy(), grid)              # device row order == reference row order
    m1, v1, m2, v2 = eng.empty((M,)), eng.empty((M,)), eng.empty((M,)), eng.empty((M,))
    eng.posterior_grid(0, 0, M, 2.0, 0.0, mean=m1, var=v1)
    eng.posterior_rows(0, eng.to_device(grid), 2.0, 0.0, mean=m2, var=v2)
    assert (m1 - m2).abs().max().item() < 1e-12 and (v1 - v2).abs().max().item() < 1e-12
    h = M // 3                                                                    # a shard starting mid-grid gives the same rows
    m3 = eng.empty((M - h,))
    eng.posterior_grid(0, h, M - h, 2.0, 0.0, mean=m3)
    assert torch.equal(m3, m1[h:])
    eng.close()


@pytest.mark.parametrize("N,d,n", [(80, 2, 120), (257, 2, 90), (300, 2, 100), (330, 3, 30), (384, 2, 80), (512, 2, 70)])
def test_multi_output_launch_equals_single_launches(N, d, n):
    """Three GPs on one factorisation (so_fit_like) evaluated by ONE launch -- one contraction, one V.z per GP -- against three
    single launches, on the grid path (every tile plan: 48-row double buffer, six block rows per warp, ring) and on explicit
    rows: means, variances, bounds and the AND-ed safe bit are bit-identical."""
    rs = np.random.RandomState(N)
    X = rs.uniform(-2, 2, (N, d))
    Ys = [np.sin(X.sum(1) + o) + 0.05 * rs.randn(N) for o in (0.0, 0.7, 1.9)]
    ls = rs.uniform(0.7, 1.3, d)
    grid = sb.linearly_spaced_combinations([(-2.0, 2.0)] * d, n)
    M = grid.shape[0]
    from safeopt_b200.utilities import detect_grid
    eng = DeviceEngine(max_gps=3)
    eng.fit(0, X, Ys[0], 0, ls, 2.0, 0.05 ** 2)
    eng.fit_like(1, 0, Ys[1])
    eng.fit_like(2, 0, Ys[2])
    eng.define_grid(detect_grid(grid))
    eng.prepare_grid(0, 0, M)
    fmins = [-0.2, 0.0, 0.1]
    for rows in (None, eng.to_device(grid)):
        out = {}
        for multi in (True, False):
            means = [eng.empty((M,)) for _ in range(3)]
            variances = [eng.empty((M,)) for _ in range(3)]
            Q, S = eng.empty((M, 6)), eng.zeros((M,), "u8")
            if multi:
                assert eng.posterior_multi([0, 1, 2], rows, 0, M, 2.0, fmins, means=means, variances=variances, Q=Q, q_cols=[0, 2, 4], S=S,
                                           safe_mode=_lib.SAFE_WRITE)
            else:
                for i in range(3):
                    mode = _lib.SAFE_WRITE if i == 0 else _lib.SAFE_AND
                    if rows is None:
                        if i:
                            eng.prepare_grid(i, 0, M)
                        eng.posterior_grid(i, 0, M, 2.0, fmins[i], mean=means[i], var=variances[i], Q=Q, q_col=2 * i, S=S, safe_mode=mode)
                    else:
                        eng.posterior_rows(i, rows, 2.0, fmins[i], mean=means[i], var=variances[i], Q=Q, q_col=2 * i, S=S, safe_mode=mode)
            out[multi] = [t.cpu().numpy() for t in means + variances + [Q, S]]
        for a, b in zip(out[True], out[False]):
            assert np.array_equal(a, b)
        assert 0 < out[True][-1].sum() < M                     # the safe bit discriminates on this problem
    eng.close()


@pytest.mark.parametrize("explicit", [False, True])
def test_shared_factorisation_equals_separate_launches(explicit, monkeypatch):
    """GPs with identical inputs, kernel and noise are evaluated by one launch (one contraction, one V.z per GP):
    bit-identical to evaluating them one by one, on the grid path and on explicit rows."""
    if explicit:
        monkeypatch.setenv("SAFEOPT_B200_GRID_FAST_PATH", "0")
    g = load_golden("config_C3_n80")                      # three GPs, same kernel, same X, different Y
    res = {}
    for share in ("1", "0"):
        monkeypatch.setenv("SAFEOPT_B200_SHARE_FITS", share)
        gps, grid, fmin = golden_problem(g, "gpu")
        opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
        opt.optimize()
        assert opt._fits.groups == ([[0, 1, 2]] if share == "1" else [[0], [1], [2]])
        res[share] = (opt.Q.copy(), opt.S.copy(), opt.M.copy(), opt.last_query_row, opt._mean_d.cpu().numpy(), opt._var_d.cpu().numpy(),
                      opt._engine.launches)
    for a, b in zip(res["1"][:6], res["0"][:6]):
        assert np.array_equal(a, b)
    assert res["1"][6] < res["0"][6]
    # a GP with its own kernel leaves the group
    gps, grid, fmin = golden_problem(g, "gpu")
    gps[1].kern.variance = np.array([1.7])
    monkeypatch.setenv("SAFEOPT_B200_SHARE_FITS", "1")
    opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
    opt.update_confidence_intervals()
    assert opt._fits.groups == [[0, 2], [1]]


def test_shared_factorisation_swarm_fitness(monkeypatch):
    gl = load_golden("swarm_fitness_3d")
    X, Y = gl["X"], gl["Y"]
    d = X.shape[1]
    out = {}
    for share in ("1", "0"):
        monkeypatch.setenv("SAFEOPT_B200_SHARE_FITS", share)
        gps = [sb.GPRegression(X, Y[:, [i]], kernel=sb.RBF(d, variance=2.0, lengthscale=np.ones(d), ARD=True), noise_var=float(gl["noise_var"]))
               for i in range(Y.shape[1])]
        opt = sb.SafeOptSwarm(gps, list(gl["fmin"]), bounds=[(-1.0, 1.0)] * d, beta=float(gl["beta"]), swarm_size=20)
        opt.best_lower_bound = float(gl["best_lower_bound"])
        out[share] = [opt._compute_particle_fitness(kind, gl["particles"]) for kind in ["greedy", "maximizers", "expanders", "safe_set"]]
        assert opt._fits.groups == ([[0, 1]] if share == "1" else [[0], [1]])
    for (va, sa), (vb, sb_) in zip(out["1"], out["0"]):
        assert np.array_equal(va, vb) and np.array_equal(sa, sb_)


def test_grid_tables_for_a_row_block_only():
    """so_grid_prepare_rows: a rank builds the scaled-operand table for its own row block; rows inside give the bits of a
    full preparation, rows outside are refused."""
    from safeopt_b200.utilities import detect_grid
    from safeopt_b200._lib import DeviceError
    w = workloads.grid_workload("t", 4, [9, 14, 40, 11], 40)          # 3960 fast rows x 14 slow indices
    grid = sb.linearly_spaced_combinations(w.bounds, [9, 14, 40, 11])
    M = grid.shape[0]
    eng = DeviceEngine(max_gps=1)
    eng.fit(0, w.X, w.Y[:, 0], 0, w.lengthscale, w.variance, w.noise_var)
    eng.define_grid(detect_grid(grid))
    eng.prepare_grid(0)
    full_m, full_v = eng.empty((M,)), eng.empty((M,))
    eng.posterior_grid(0, 0, M, 2.0, 0.0, mean=full_m, var=full_v)
    lo, hi = 2 * M // 5 + 3, 4 * M // 5 + 1
    eng.prepare_grid(0, lo, hi - lo)
    m, v = eng.empty((hi - lo,)), eng.empty((hi - lo,))
    eng.posterior_grid(0, lo, hi - lo, 2.0, 0.0, mean=m, var=v)
    assert torch.equal(m, full_m[lo:hi]) and torch.equal(v, full_v[lo:hi])
    with pytest.raises(DeviceError):
        eng.posterior_grid(0, 0, M, 2.0, 0.0, mean=full_m, var=full_v)
    eng.close()


def test_chained_set_passes_equal_host_chained():
    """The *_chain entry points (scalars read from device records) give the records and masks of the host-scalar ones."""
    from safeopt_b200.engine import MAX_REC_DTYPE, SAFE_REC_DTYPE
    g = load_golden("expander_g2")
    gps, grid, fmin = golden_problem(g, "gpu")
    opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
    opt.update_confidence_intervals()
    eng, G, M = opt._engine, len(gps), grid.shape[0]
    rec_s, rec_m = eng.zeros((1, 64), "u8"), eng.zeros((1, 64), "u8")
    eng.reduce_safe(opt._Q_d, G, 0, opt._S_d, rec_s)
    safe = rec_s.cpu().numpy().view(SAFE_REC_DTYPE).reshape(-1)[0]
    thr = np.full(G, float(g["threshold"]) * float(g["beta"]))
    out = {}
    for chained in (False, True):
        Mm, key, row, cnt = eng.zeros((M,), "u8"), eng.empty((M,)), eng.empty((M,), "i64"), eng.zeros((1,), "i64")
        if chained:
            eng.maximizers_chain(opt._Q_d, G, 0, opt._S_d, rec_s, 1, opt.scaling, Mm, rec_m)
        else:
            eng.maximizers(opt._Q_d, G, 0, opt._S_d, float(safe["max_l0"]), opt.scaling, Mm, rec_m)
        mx = rec_m.cpu().numpy().view(MAX_REC_DTYPE).reshape(-1)[0].copy()
        if chained:
            eng.candidates_chain(opt._Q_d, G, 0, opt._S_d, Mm, rec_m, 1, opt.scaling, thr, None, key, row, cnt)
        else:
            eng.candidates(opt._Q_d, G, 0, opt._S_d, Mm, float(mx["max_width0"]) / opt.scaling[0], opt.scaling, thr, None, key, row, cnt)
        n = int(cnt.item())
        order = torch.argsort(row[:n])
        out[chained] = (mx, Mm.cpu().numpy(), row[:n][order].cpu().numpy(), key[:n][order].cpu().numpy())
    assert out[False][0] == out[True][0] and int(out[True][0]["n_max"]) == int(unpack_mask(g["M"], M).sum())
    for a, b in zip(out[False][1:], out[True][1:]):
        assert np.array_equal(a, b)
    assert out[True][2].size > 0


def _check_argmax(row, row_ref, values_ref, mask_ref, tol):
    """`row` must be the reference's argmax row; if the reference's values tie within 100*tol, any tied row passes."""
    cand = np.flatnonzero(mask_ref)
    top = np.sort(values_ref[cand])[::-1]
    gap = top[0] - top[1] if top.size > 1 else np.inf
    if gap > 100 * tol:
        assert row == row_ref
    else:
        assert mask_ref[row] and abs(values_ref[row] - values_ref[row_ref]) < tol


# ---------------------------------------------------------------- golden fixtures produced by the reference
@pytest.mark.parametrize("explicit", [False, True])
@pytest.mark.parametrize("name", GRID_CASES)
def test_safeopt_matches_golden(name, explicit, monkeypatch):
    g = load_golden(name)
    n_rows = int(g["n_rows"])
    gps, grid, fmin = golden_problem(g, "gpu")
    if explicit:
        monkeypatch.setenv("SAFEOPT_B200_GRID_FAST_PATH", "0")       # explicit-rows kernels on the same parameter set
    opt = sb.SafeOpt(gps if len(gps) > 1 else gps[0], grid, fmin if len(fmin) > 1 else fmin[0], beta=float(g["beta"]),
                     threshold=float(g["threshold"]), lipschitz=golden_lipschitz(g))
    assert (opt._grid_axes is None) == explicit
    if bool(g["full_sets"]):
        opt.update_confidence_intervals()
        opt.compute_sets(full_sets=True)
        x = opt.get_new_query_point()
    else:
        x = opt.optimize()
    tol = 1e-9 * 2 * np.sqrt(float(g["variance"]))
    assert np.abs(opt.Q - g["Q"]).max() < tol
    assert min(float(g["margin_S"]), float(g["margin_M"])) > 100 * tol          # masks are decidable at this tolerance
    assert np.array_equal(opt.S, unpack_mask(g["S"], n_rows))
    assert np.array_equal(opt.M, unpack_mask(g["M"], n_rows))
    assert np.array_equal(opt.G, unpack_mask(g["G"], n_rows))
    # argmax rows: identical whenever the reference's own best/second-best gap is decidable at this tolerance; the 1-D
    # doctest fixture is mirror symmetric (rows 27 and 72 tie EXACTLY in the reference, which then takes the first), so
    # there any row whose reference value ties with the reference's pick within `tol` is accepted
    S_ref, M_ref, G_ref = (unpack_mask(g[k], n_rows) for k in "SMG")
    width = ((g["Q"][:, 1::2] - g["Q"][:, ::2]) / opt.scaling).max(axis=1)
    _check_argmax(opt.last_query_row, int(g["row_next"]), width, M_ref | G_ref, tol)
    assert np.array_equal(x, grid[opt.last_query_row])
    mx = opt.get_maximum()
    assert abs(mx[1] - float(g["max_val"])) < tol
    lower0 = g["Q"][:, 0]
    _check_argmax(int(np.flatnonzero((grid == mx[0]).all(axis=1))[0]), int(np.flatnonzero((grid == g["max_x"]).all(axis=1))[0]),
                  lower0, S_ref, tol)
    opt.optimize(ucb=True)
    _check_argmax(opt.last_query_row, int(g["row_ucb"]), g["Q"][:, 1], S_ref, tol)


@pytest.mark.parametrize("name", ["context_1p1c", "context_1p1c_lipschitz"])
def test_contexts_match_golden(name):
    """Contexts (gp_opt.py:424-451) with a product of RBF kernels on disjoint dims (examples/context_example.ipynb)."""
    g = load_golden(name)
    kern = sb.RBF(1, variance=float(g["var0"]), lengthscale=float(g["ls0"]), active_dims=[0]) * \
        sb.RBF(1, variance=float(g["var1"]), lengthscale=float(g["ls1"]), active_dims=[1])
    gp = sb.GPRegression(g["X"], g["Y"], kernel=kern, noise_var=float(g["noise_var"]))
    lip = g["lipschitz"]
    opt = sb.SafeOpt(gp, g["pset"], float(g["fmin"]), num_contexts=1, beta=float(g["beta"]), threshold=float(g["threshold"]),
                     lipschitz=None if lip.size == 0 else float(lip[0]))
    n_rows = int(g["n_rows"])
    # the parameter set is a product grid: contexts ride along as one-point axes and the separable-table grid kernels are
    # used (no M x d candidate array on the device)
    assert opt._grid_axes is not None and len(opt._grid_axes) == 2 and opt._rows_d is None
    with pytest.raises(ValueError):
        opt.optimize()                                   # a context is required (gp_opt.py:448-450)
    for k in range(2):
        ctx = np.array([float(g["ctx%d" % k])])
        x = opt.optimize(context=ctx)
        assert np.abs(opt.Q - g["Q%d" % k]).max() < 1e-9 * 2 * np.sqrt(3.0)
        assert np.array_equal(opt.S, unpack_mask(g["S%d" % k], n_rows))
        assert np.array_equal(opt.M, unpack_mask(g["M%d" % k], n_rows))
        assert np.array_equal(opt.G, unpack_mask(g["G%d" % k], n_rows))
        assert np.array_equal(x, g["x%d" % k]) and x.shape == (1,)
        assert np.array_equal(opt.context, ctx) and opt._grid_axes[1][0] == ctx[0] and opt._rows_d is None
        mx = opt.get_maximum(context=ctx)
        assert np.array_equal(mx[0], g["maxx%d" % k]) and abs(mx[1] - float(g["maxv%d" % k])) < 1e-9
    opt.add_new_data_point(x, np.array([[0.5]]), context=ctx)
    assert opt.x.shape == (13, 2) and opt.x[-1, 1] == ctx[0]


def test_use_lipschitz_switch():
    gp = sb.GPRegression(np.array([[0.0]]), np.array([[1.0]]), noise_var=0.01 ** 2)
    opt = sb.SafeOpt(gp, sb.linearly_spaced_combinations([(-1, 1)], 50), fmin=0.0)
    assert opt.use_lipschitz is False
    with pytest.raises(ValueError):
        opt.use_lipschitz = True                          # gp_opt.py:403-407


def test_bo_loop_matches_golden():
    g = load_golden("bo_loop_2d")
    gp = sb.GPRegression(g["X"], g["Y"], kernel=sb.RBF(2, variance=2.0, lengthscale=np.ones(2), ARD=True), noise_var=float(g["noise_var"]))
    grid = sb.linearly_spaced_combinations([tuple(b) for b in g["bounds"]], int(g["num_samples"]))
    opt = sb.SafeOpt(gp, grid, float(g["fmin"]), beta=float(g["beta"]), threshold=float(g["threshold"]))
    for it, row_ref in enumerate(g["rows"]):
        x = opt.optimize()
        assert opt.last_query_row == int(row_ref), "trajectory diverged at iteration %d" % it
        assert int(opt.G.sum()) == int(g["n_expanders"][it])
        opt.add_new_data_point(x, np.array([[g["ys"][it]]]))
    # one from-scratch fit, then one-point appends (f4) -- the trajectory above is the parity check of both
    assert opt.t == 6 + len(g["rows"]) and opt._fits.refits + opt._fits.appends == len(g["rows"]) and opt._fits.appends >= 15


def test_no_safe_points_raises_like_reference():
    gp = sb.GPRegression(np.array([[0.0]]), np.array([[-1.0]]), noise_var=0.01 ** 2)
    opt = sb.SafeOpt(gp, sb.linearly_spaced_combinations([(-1, 1)], 50), fmin=0.0)
    with pytest.raises(EnvironmentError):
        opt.optimize()
    assert opt.get_maximum() is None
    assert not opt.S.any() and not opt.M.any() and not opt.G.any()


# ---------------------------------------------------------------- swarm
def test_swarm_fitness_matches_golden():
    g = load_golden("swarm_fitness_3d")
    X, Y = g["X"], g["Y"]
    d = X.shape[1]
    gps = [sb.GPRegression(X, Y[:, [i]], kernel=sb.RBF(d, variance=2.0, lengthscale=np.ones(d), ARD=True), noise_var=float(g["noise_var"]))
           for i in range(Y.shape[1])]
    opt = sb.SafeOptSwarm(gps, list(g["fmin"]), bounds=[(-1.0, 1.0)] * d, beta=float(g["beta"]), swarm_size=20)
    assert np.allclose(opt.optimal_velocities, g["velocities"], rtol=0, atol=1e-12)
    assert np.array_equal(opt._compute_penalty(g["penalty_in"].copy()), g["penalty_out"])
    opt.best_lower_bound = float(g["best_lower_bound"])
    for kind in ["greedy", "maximizers", "expanders", "safe_set"]:
        v, s = opt._compute_particle_fitness(kind, g["particles"])
        ref = g["values_" + kind]
        assert np.abs(v - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), kind
        assert np.array_equal(s, g["safe_" + kind]), kind
    with pytest.raises(AssertionError):
        opt._compute_particle_fitness("bogus", g["particles"])


def test_swarm_step_kernels_match_port():
    rs = np.random.RandomState(4)
    P, d = 777, 3
    eng = DeviceEngine(max_gps=1)
    pos, vel = rs.uniform(-1, 1, (P, d)), rs.uniform(0, 0.1, (P, d))
    bpos, bval = rs.uniform(-1, 1, (P, d)), rs.randn(P)
    gbest = bpos[np.argmax(bval)].copy()
    r = rs.rand(2 * P, d)
    vs = np.array([0.1, 0.2, 0.15])
    bounds = np.array([(-1.0, 1.0)] * d)
    values, safe = rs.randn(P), rs.rand(P) > 0.3
    fit = lambda x: (values, safe)
    p2, v2, bp2, bv2, gb2 = port.pso_step(pos, vel, bpos, bval, gbest, r[:P], r[P:], 0.73, vs, bounds, fit)
    pd, vd, bpd, bvd = [eng.to_device(a.copy()) for a in (pos, vel, bpos, bval)]
    eng.swarm_step(pd, vd, bpd, eng.to_device(gbest), eng.to_device(r), 0.73, vs, bounds)
    assert np.array_equal(pd.cpu().numpy(), p2) and np.array_equal(vd.cpu().numpy(), v2)
    idx = eng.zeros((1,), "i64")
    eng.swarm_update_best(pd, eng.to_device(values), eng.to_device(safe.astype(np.uint8)), bpd, bvd, idx)
    assert np.array_equal(bvd.cpu().numpy(), bv2) and np.array_equal(bpd.cpu().numpy(), bp2)
    assert int(idx.item()) == int(np.argmax(bv2))
    eng.close()


def test_safeoptswarm_runs_and_reference_smoke_test():
    # /root/reference/safeopt/tests/test_swarm.py:13-22
    gp = sb.GPRegression(np.array([[0.0]]), np.array([[-1.0]]), noise_var=0.01 ** 2)
    opt = sb.SafeOptSwarm(gp, fmin=[0.0], bounds=[[-1.0, 1.0]])
    with pytest.raises(RuntimeError):
        opt.optimize()
    # docstring example gp_opt.py:755-777
    gp = sb.GPRegression(np.array([[0.0]]), np.array([[1.0]]), noise_var=0.01 ** 2)
    opt = sb.SafeOptSwarm(gp, fmin=[0.0], bounds=[[-1.0, 1.0]])
    np.random.seed(0)
    x = opt.optimize()
    assert x.shape == (1,) and -1.0 <= x[0] <= 1.0
    opt.add_new_data_point(x, np.array([[1.0]]))
    assert opt.t == 2 and opt.get_maximum()[1] == 1.0


def test_device_swarm_large_matches_host_swarm_logic():
    w = workloads.swarm_workload(n_particles=4096, n_train=64)
    gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
           for i in range(w.n_gps)]
    opt = sb.SafeOptSwarm(gps, [0.0, 0.2], bounds=w.bounds, beta=2.0)
    opt.best_lower_bound = 0.5
    fit_dev = lambda p: opt._fitness_device("maximizers", p)
    fit_host = lambda p: opt._compute_particle_fitness("maximizers", p)
    np.random.seed(3)
    hs = sb.SwarmOptimization(w.n_particles, opt.optimal_velocities, fit_host, bounds=w.bounds)
    hs.init_swarm(w.particles.copy())
    hs.run_swarm(5)
    np.random.seed(3)
    ds = sb.DeviceSwarm(opt._engine, opt.optimal_velocities, fit_dev, bounds=w.bounds, rng="host")
    ds.init_swarm(w.particles.copy())
    ds.run_swarm(5)
    assert np.abs(ds.positions.cpu().numpy() - hs.positions).max() < 1e-9
    assert np.abs(ds.best_values.cpu().numpy() - hs.best_values).max() < 1e-8
    assert np.abs(ds.global_best - hs.global_best).max() < 1e-9


def _swarm_query_problem(g, **kw):
    X, Y = g["X"], g["Y"]
    d = X.shape[1]
    cls = {0: sb.RBF, 1: sb.Matern32, 2: sb.Matern52}[int(g["kind"])]
    gps = [sb.GPRegression(X, Y[:, [i]], kernel=cls(d, variance=float(g["variance"]), lengthscale=g["lengthscale"], ARD=True),
                           noise_var=float(g["noise_var"])) for i in range(Y.shape[1])]
    opt = sb.SafeOptSwarm(gps, list(g["fmin"]), bounds=[tuple(b) for b in g["bounds"]], beta=float(g["beta"]),
                          swarm_size=int(g["swarm_size"]), **kw)
    opt.max_iters = int(g["max_iters"])
    return opt


@pytest.mark.parametrize("name", ["swarm_query_2d", "swarm_query_2d_mat32"])
def test_safeset_insertion_matches_golden(name):
    """f1: the device insertion kernels on the reference's own swarms (gp_opt.py:1088-1110)."""
    g = load_golden(name)
    opt = _swarm_query_problem(g)
    for stage in ["maximizers", "expanders"]:
        before, after = g[stage + "_S_before"], g[stage + "_S_after"]
        opt.S = before.copy()
        new = opt._select_new_safe_points(g[stage + "_best_positions"])
        assert np.array_equal(new, after[before.shape[0]:]), stage


@pytest.mark.parametrize("kind,d,n,m", [(0, 3, 3000, 500), (1, 6, 2500, 0), (2, 2, 1024, 300), (0, 16, 1500, 40), (0, 1, 5, 3)])
def test_safeset_kernels_match_port(kind, d, n, m):
    """Filter + blocked sequential walk vs the port on random candidates spanning several 1024-blocks."""
    rs = np.random.RandomState(100 + d)
    ls = rs.uniform(0.5, 1.5, d)
    X, Y = rs.uniform(-1, 1, (8, d)), rs.randn(8)
    eng = DeviceEngine(max_gps=1)
    eng.fit(0, X, Y, kind, ls, 1.7, 0.01)
    kern = oracle_kernel(kind, d, 1.7, ls)
    spread = {1: 8.0, 2: 6.0, 3: 2.5, 6: 0.9, 16: 0.16}[d]        # dense enough that a good share is rejected
    cand, ref = rs.uniform(-spread, spread, (n, d)), rs.uniform(-spread, spread, (m, d))
    scale0 = np.sqrt(1.7)
    acc_ref, margin = port.select_new_safe_points(kern, ref, cand, scale0)
    assert margin > 1e-9
    cand_d = eng.to_device(cand)
    keep = eng.empty((n,), "u8")
    eng.safeset_filter(0, cand_d, eng.to_device(ref) if m else eng.empty((0, d)), scale0 ** 2, 0.95, keep)
    if m:
        keep_ref = np.all(kern.K(cand, ref) / scale0 ** 2 <= 0.95, axis=1)
        assert np.array_equal(keep.cpu().numpy().astype(bool), keep_ref)
    accept, pos, cnt = eng.empty((n,), "u8"), eng.empty((n, d)), eng.zeros((1,), "i64")
    eng.safeset_insert(0, cand_d, keep, scale0 ** 2, 0.95, accept, pos, cnt)
    acc = accept.cpu().numpy().astype(bool)
    assert 0 < acc_ref.sum() < n or n < 10
    assert np.array_equal(acc, acc_ref)
    assert int(cnt.item()) == acc_ref.sum() and np.array_equal(pos[:int(cnt.item())].cpu().numpy(), cand[acc_ref])
    eng.close()


@pytest.mark.parametrize("name", ["swarm_query_2d", "swarm_query_2d_mat32"])
@pytest.mark.parametrize("backend", ["host", "device"])
def test_safeoptswarm_trajectory_matches_golden(name, backend):
    """SafeOptSwarm.optimize() unrolled like gp_opt.py:1136-1177 with the reference's random stream: every
    swarm's best positions, the safe-set growth and the returned points follow the reference."""
    g = load_golden(name)
    opt = _swarm_query_problem(g, swarm_backend=backend, rng="host")
    np.random.seed(int(g["seed"]))
    for stage in ["greedy", "maximizers", "expanders"]:
        assert np.array_equal(opt.S, g[stage + "_S_before"]), stage
        x, v = opt.get_new_query_point(stage)
        if stage == "greedy":
            opt.greedy, opt.best_lower_bound = x, v
        sw = opt.swarms[stage]
        bp = sw.best_positions.cpu().numpy() if backend == "device" else sw.best_positions
        assert np.abs(bp - g[stage + "_best_positions"]).max() < 1e-7, stage
        assert np.abs(np.asarray(x) - g[stage + "_x"]).max() < 1e-7 and np.abs(np.asarray(v) - g[stage + "_v"]).max() < 1e-7, stage
        assert opt.S.shape == g[stage + "_S_after"].shape and np.abs(opt.S - g[stage + "_S_after"]).max() < 1e-7, stage
    opt2 = _swarm_query_problem(g, swarm_backend=backend, rng="host")
    np.random.seed(int(g["seed"]))
    assert np.abs(opt2.optimize() - g["x_next"]).max() < 1e-7
    assert opt2.S.shape == g["S_final"].shape


def test_device_swarm_device_rng_runs():
    g = load_golden("swarm_query_2d")
    opt = _swarm_query_problem(g, swarm_backend="device", rng="device", seed=3)
    opt.swarm_size = 4096
    np.random.seed(0)
    x = opt.optimize()
    assert x.shape == (2,) and np.all(np.abs(x) <= 1.5)
    _, safe = opt._compute_particle_fitness("safe_set", opt.S)
    assert opt.S.shape[0] > g["X"].shape[0]
    # every pair of points in the grown safe set respects the correlation limit (gp_opt.py:1105)
    kern = oracle_kernel(0, 2, float(g["variance"]), g["lengthscale"])
    C = kern.K(opt.S[g["X"].shape[0]:], opt.S) / float(g["variance"])
    np.fill_diagonal(C[:, g["X"].shape[0]:], 0.0)
    assert C.max() <= 0.95 + 1e-12


# ---------------------------------------------------------------- full-size properties (BASELINE config 4 and 2)
def test_config_c4_full_size_properties():
    w = workloads.config("C4")
    grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
    gp = sb.GPRegression(w.X, w.Y, kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
    opt = sb.SafeOpt(gp, grid, 0.0, beta=w.beta, threshold=w.threshold)
    assert opt._grid_axes is not None and grid.shape[0] == 6_250_000
    x = opt.optimize()
    Q, S, M, G = opt.Q, opt.S, opt.M, opt.G
    # idempotence of the set logic on the device's own Q (bit-exact, size independent)
    assert np.array_equal(S, Q[:, 0] > 0.0)
    best_l = Q[S, 0].max()
    assert np.array_equal(M, S & (Q[:, 1] >= best_l))
    MG = M | G
    width = (Q[:, 1] - Q[:, 0]) / opt.scaling[0]
    assert opt.last_query_row == np.flatnonzero(MG)[np.argmax(width[MG])]
    assert np.array_equal(x, grid[opt.last_query_row])
    # EVERY row against the oracle port (6.25e6 rows in 250k-row chunks, ~30 s of host time): Q within tolerance, the masks
    # S / M / G and the query row bit-exact -- BASELINE's "safe-set masks bit-exact vs reference" at d=4, N=256, full size
    go = gpy_lite.GPRegression(w.X, w.Y, kernel=gpy_lite.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
    Qo = port.confidence_intervals([go], grid, w.beta, chunk=250_000)
    assert np.abs(Q - Qo).max() < 1e-9 * 2 * np.sqrt(w.variance)
    fm = np.array([0.0])
    trace = {}
    So, Mo, Go = port.compute_sets([go], grid, Qo, fm, w.beta, opt.scaling, w.threshold, trace=trace)
    margin_S = np.abs(Qo[:, 0] - 0.0).min()
    margin_M = np.abs(Qo[So, 1] - Qo[So, 0].max()).min()
    print("C4 full size: |dQ| = %.2e, min margin S = %.3e, M = %.3e, |S| = %d, |M| = %d, |G| = %d, candidates = %d" % (
        np.abs(Q - Qo).max(), margin_S, margin_M, So.sum(), Mo.sum(), Go.sum(), trace.get("n_candidates", 0)))
    assert margin_S > 1e-8 and margin_M > 1e-8, "workload sits on a knife edge: masks would not be comparable"
    assert np.array_equal(S, So) and np.array_equal(M, Mo) and np.array_equal(G, Go)
    _, row_o = port.new_query_point(grid, Qo, So, Mo, Go, opt.scaling)
    assert opt.last_query_row == row_o
    # sharding invariance: the second half computed as its own shard is bit-identical
    eng = opt._engine
    h = grid.shape[0] // 2
    m_half = eng.empty((grid.shape[0] - h,))
    eng.posterior_grid(0, h, grid.shape[0] - h, w.beta, 0.0, mean=m_half)
    assert torch.equal(m_half, opt._mean_d[0, h:])
    # ucb / maximum agree with NumPy on the device's Q
    mx = opt.get_maximum()
    assert mx[1] == Q[S, 0].max() and np.array_equal(mx[0], grid[np.flatnonzero(S)[np.argmax(Q[S, 0])]])


def test_config_c3_three_gps_properties():
    w = workloads.config("C3")
    grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
    gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
           for i in range(3)]
    opt = sb.SafeOpt(gps, grid, w.fmin, beta=w.beta, threshold=w.threshold)
    opt.optimize()
    assert (opt._fits.refits, opt._fits.copies) == (3, 2)      # one factorisation, two device-to-device copies (so_fit_like)
    Q, S, M = opt.Q, opt.S, opt.M
    assert Q.shape == (250000, 6)
    assert np.array_equal(S, np.all(Q[:, ::2] > 0.0, axis=1))
    assert np.array_equal(M, S & (Q[:, 1] >= Q[S, 0].max()))
    # every row and every GP against the oracle port (250k rows x 3 GPs): Q within tolerance, masks and query row bit-exact
    gos = [gpy_lite.GPRegression(w.X, w.Y[:, [i]], kernel=gpy_lite.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True),
                                 noise_var=w.noise_var) for i in range(3)]
    Qo = port.confidence_intervals(gos, grid, w.beta, chunk=125_000)
    assert np.abs(Q - Qo).max() < 1e-9 * 2 * np.sqrt(w.variance)
    So, Mo, Go = port.compute_sets(gos, grid, Qo, np.asarray(w.fmin, dtype=float), w.beta, opt.scaling, w.threshold)
    margin_S = np.abs(Qo[:, ::2] - 0.0).min()
    margin_M = np.abs(Qo[So, 1] - Qo[So, 0].max()).min()
    print("C3 full size: |dQ| = %.2e, min margin S = %.3e, M = %.3e" % (np.abs(Q - Qo).max(), margin_S, margin_M))
    assert margin_S > 1e-8 and margin_M > 1e-8
    assert np.array_equal(S, So) and np.array_equal(M, Mo) and np.array_equal(opt.G, Go)
    assert opt.last_query_row == port.new_query_point(grid, Qo, So, Mo, Go, opt.scaling)[1]


# ---------------------------------------------------------------- the CPU stand-in engine honours the same contracts
def test_stand_in_engine_agrees_with_the_device():
    """tests/fake_engine.py (NumPy, used by the CPU orchestration tests) and the real engine give the same answers to the
    same calls: posterior + bounds + safe bits, the three chained set passes, both expander tests."""
    from fake_engine import FakeEngine
    from safeopt_b200.engine import MAX_REC_DTYPE, SAFE_REC_DTYPE
    from safeopt_b200.utilities import detect_grid
    g = load_golden("expander_g2")
    X, Y = g["X"], g["Y"]
    grid = port.linearly_spaced_combinations([tuple(b) for b in g["bounds"]], [int(v) for v in np.atleast_1d(g["num_samples"])])
    M, G, beta = grid.shape[0], Y.shape[1], float(g["beta"])
    fmin = [float(v) for v in g["fmin"]]
    scaling = np.full(G, np.sqrt(float(g["variance"])))
    thr = np.full(G, float(g["threshold"]) * beta)
    out = {}
    for name, eng in (("dev", DeviceEngine(max_gps=G)), ("fake", FakeEngine(max_gps=G))):
        eng.define_grid(detect_grid(grid))
        for i in range(G):
            eng.fit(i, X, Y[:, i], int(g["kind"]), g["lengthscale"], float(g["variance"]), float(g["noise_var"]))
            eng.prepare_grid(i)
        Q, S, Mm = eng.empty((M, 2 * G)), eng.zeros((M,), "u8"), eng.zeros((M,), "u8")
        mean, var = eng.empty((G, M)), eng.empty((G, M))
        for i in range(G):
            eng.posterior_grid(i, 0, M, beta, fmin[i], mean=mean[i], var=var[i], Q=Q, q_col=2 * i, S=S,
                               safe_mode=_lib.SAFE_WRITE if i == 0 else _lib.SAFE_AND)
        rs, rm = eng.zeros((1, 64), "u8"), eng.zeros((1, 64), "u8")
        key, row, cnt = eng.empty((M,)), eng.empty((M,), "i64"), eng.zeros((1,), "i64")
        eng.reduce_safe(Q, G, 0, S, rs)
        eng.maximizers_chain(Q, G, 0, S, rs, 1, scaling, Mm, rm)
        eng.candidates_chain(Q, G, 0, S, Mm, rm, 1, scaling, thr, None, key, row, cnt)
        n = int(cnt.cpu().numpy()[0])
        rows = np.sort(row.cpu().numpy()[:n])
        # expander tests on the four widest candidates
        Qh, meanh, varh = Q.cpu().numpy(), mean.cpu().numpy(), var.cpu().numpy()
        cand = rows[np.argsort(-(Qh[rows, 1] - Qh[rows, 0]))[:4]]
        xc = eng.to_device(grid[cand])
        flags_gp, flags_lip = eng.zeros((32,), "u8"), eng.zeros((32,), "u8")
        eng.expander_check(0, None, 0, M, S, mean[0], var[0], xc, eng.to_device(meanh[0, cand]), eng.to_device(varh[0, cand]),
                           eng.to_device(Qh[cand, 1]), beta, fmin[0], flags_gp)
        eng.expander_lipschitz(None, 2, 0, M, S, xc, eng.to_device(Qh[cand, 1]), 2.0, fmin[0], flags_lip)
        out[name] = dict(Q=Qh, S=S.cpu().numpy(), M=Mm.cpu().numpy(), safe=rs.cpu().numpy().view(SAFE_REC_DTYPE).reshape(-1)[0],
                         mx=rm.cpu().numpy().view(MAX_REC_DTYPE).reshape(-1)[0], rows=rows, cand=cand,
                         fgp=flags_gp.cpu().numpy()[:4].copy(), flip=flags_lip.cpu().numpy()[:4].copy())
        eng.close()
    d, f = out["dev"], out["fake"]
    assert np.abs(d["Q"] - f["Q"]).max() < 1e-9 * 2 * np.sqrt(2.0)
    assert np.array_equal(d["S"], f["S"]) and np.array_equal(d["M"], f["M"]) and np.array_equal(d["rows"], f["rows"])
    assert np.array_equal(d["cand"], f["cand"]) and np.array_equal(d["fgp"], f["fgp"]) and np.array_equal(d["flip"], f["flip"])
    for k in ("n_safe", "argmax_l0", "argmax_u0"):
        assert d["safe"][k] == f["safe"][k]
    assert abs(d["safe"]["max_l0"] - f["safe"]["max_l0"]) < 1e-9 and d["mx"]["n_max"] == f["mx"]["n_max"]
    assert d["mx"]["best_row"] == f["mx"]["best_row"] and abs(d["mx"]["max_width0"] - f["mx"]["max_width0"]) < 1e-9


# ---------------------------------------------------------------- fused set pass, in-kernel exchange, graph replay
@pytest.mark.parametrize("name", ["doctest_1d", "config_C2", "expander_g1", "expander_g2", "expander_tight", "full_sets_g1", "lipschitz_g2"])
def test_fused_set_pass_equals_chained_passes(name, monkeypatch):
    """so_sets_fused (one cooperative launch, records exchanged through the peer buffers) against the three chained kernels."""
    g = load_golden(name)
    n_rows = int(g["n_rows"])
    full = bool(g["full_sets"]) if "full_sets" in g.files else False
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SAFEOPT_B200_FUSED_SETS", mode)
        gps, grid, fmin = golden_problem(g, "gpu")
        opt = sb.SafeOpt(gps if len(gps) > 1 else gps[0], grid, fmin if len(gps) > 1 else fmin[0], lipschitz=golden_lipschitz(g),
                         beta=float(g["beta"]), threshold=float(g["threshold"]))
        assert opt._fused == (mode == "1")
        for rep in range(3):                                  # epochs alternate the parity slots of the exchange buffer
            opt.update_confidence_intervals()
            opt.compute_sets(full_sets=full)
            got[(mode, rep)] = (opt.S.copy(), opt.M.copy(), opt.G.copy(), dict(opt._safe_info), dict(opt._max_info or {}),
                                opt.last_trace.get("n_candidates"), None if full else opt.get_new_query_point().copy())
    for rep in range(3):
        a, b = got[("1", rep)], got[("0", rep)]
        assert all(np.array_equal(a[k], b[k]) for k in range(3)) and a[3] == b[3] and a[4] == b[4] and a[5] == b[5]
        assert full or np.array_equal(a[6], b[6])
    assert np.array_equal(got[("1", 2)][0], unpack_mask(g["S"], n_rows)) and np.array_equal(got[("1", 2)][1], unpack_mask(g["M"], n_rows))
    assert np.array_equal(got[("1", 2)][2], unpack_mask(g["G"], n_rows))


def test_replayed_posterior_launches_follow_every_change():
    """update_confidence_intervals re-issues the remembered K2 launches while fits, tables, thresholds and beta are unchanged,
    and rebuilds them as soon as one of them changes (new observation, new fmin, new beta, new hyper-parameter, new context)."""
    g = load_golden("expander_g2")
    gps, grid, fmin = golden_problem(g, "gpu")
    opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
    x0 = opt.optimize()
    tape0 = opt._k2_tape
    assert tape0 is not None and len(tape0[1]) >= 1
    Q0 = opt.Q.copy()
    launches = opt._engine.launches
    x1 = opt.optimize()                                   # replay
    assert opt._k2_tape is tape0 and np.array_equal(x0, x1) and np.array_equal(opt.Q, Q0) and opt._engine.launches > launches
    opt.fmin = opt.fmin + 0.05                            # thresholds are baked into the launches
    opt.optimize()
    assert opt._k2_tape is not tape0 and np.array_equal(opt.S, np.all(opt.Q[:, ::2] > opt.fmin, axis=1))
    tape1 = opt._k2_tape
    opt.add_new_data_point(x0, np.array([[1.0, 1.0]]))    # new fit
    opt.optimize()
    assert opt._k2_tape is not tape1
    ref = sb.SafeOpt(gps, grid, list(opt.fmin), beta=float(g["beta"]), threshold=float(g["threshold"]))
    ref.optimize()
    # (opt appended one row to its factorisation, ref factorises from scratch: equal to rounding, not bit for bit)
    assert np.abs(ref.Q - opt.Q).max() < 1e-9 and np.array_equal(ref.S, opt.S) and ref.last_query_row == opt.last_query_row
    tape2 = opt._k2_tape
    opt.beta = lambda t: 3.0
    opt.optimize()
    assert opt._k2_tape is not tape2 and np.abs((opt.Q[:, 1] - opt.Q[:, 0]) / (ref.Q[:, 1] - ref.Q[:, 0]) - 1.5).max() < 1e-9


def test_fused_set_pass_with_no_safe_rows_and_ragged_sizes():
    """Edge cases of the fused kernel: nothing safe, a single row, sizes that are not multiples of the 16-row chunks."""
    eng = DeviceEngine(max_gps=1)
    rs = np.random.RandomState(5)
    for M in [1, 15, 16, 17, 1000, 70001]:
        Qh = np.sort(rs.randn(M, 2), axis=1)
        for frac in [0.0, 0.3, 1.0]:
            Sh = (rs.rand(M) < frac).astype(np.uint8)
            Q, S, Mm = eng.to_device(Qh), eng.to_device(Sh), eng.zeros((M,), "u8")
            key, row = eng.empty((M,)), eng.empty((M,), "i64")
            host = eng.sets_fused(Q, 1, 1000, S, [1.0], [0.1], True, Mm, key, row)
            from safeopt_b200.engine import MAX_REC_DTYPE, SAFE_REC_DTYPE
            sr = host[:64].view(SAFE_REC_DTYPE)[0]
            mr = host[64:128].view(MAX_REC_DTYPE)[0]
            n_c = int(host[128:136].view(np.int64)[0])
            sb_ = Sh.astype(bool)
            assert sr["n_safe"] == sb_.sum()
            if not sb_.any():
                assert sr["argmax_l0"] == -1 and mr["n_max"] == 0 and n_c == 0 and not Mm.cpu().numpy().any()
                continue
            rows = np.flatnonzero(sb_)
            assert sr["max_l0"] == Qh[sb_, 0].max() and sr["argmax_l0"] == 1000 + rows[np.argmax(Qh[sb_, 0])]
            assert sr["argmax_u0"] == 1000 + rows[np.argmax(Qh[sb_, 1])]
            Mo = sb_ & (Qh[:, 1] >= Qh[sb_, 0].max())
            assert np.array_equal(Mm.cpu().numpy().astype(bool), Mo) and mr["n_max"] == Mo.sum()
            w = Qh[:, 1] - Qh[:, 0]
            assert mr["max_width0"] == w[Mo].max() and mr["best_row"] == 1000 + np.flatnonzero(Mo)[np.argmax(w[Mo])]
            c = sb_ & ~Mo & (w > w[Mo].max()) & (w > 0.1)
            assert n_c == c.sum() and np.array_equal(np.sort(row.cpu().numpy()[:n_c]), 1000 + np.flatnonzero(c))
    eng.close()


def test_swarm_randoms_do_not_depend_on_the_sharding():
    """so_swarm_rand / so_swarm_step_dev key their randoms by the GLOBAL particle index: any shard draws what one GPU would."""
    eng = DeviceEngine(max_gps=1)
    P, d = 1000, 6
    full = eng.empty((P, d))
    eng.swarm_rand(P, d, 0, 12345, 7, full)
    part = eng.empty((300, d))
    eng.swarm_rand(300, d, 450, 12345, 7, part)
    assert torch.equal(part, full[450:750])
    f = full.cpu().numpy()
    assert 0.0 <= f.min() and f.max() < 1.0 and abs(f.mean() - 0.5) < 0.02 and abs(f.var() - 1 / 12) < 0.01
    other = eng.empty((P, d))
    eng.swarm_rand(P, d, 0, 12345, 8, other)
    assert not torch.equal(other, full) and abs(np.corrcoef(f.ravel(), other.cpu().numpy().ravel())[0, 1]) < 0.05
    eng.close()


def test_device_swarm_graph_replay_equals_kernel_by_kernel(monkeypatch):
    """rng='device': a PSO run replayed from the captured CUDA graph is bit-identical to launching kernel by kernel, and a
    second optimise() (new greedy bound => re-capture) still is."""
    g = load_golden("swarm_query_2d")
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SAFEOPT_B200_SWARM_GRAPH", mode)
        opt = _swarm_query_problem(g, swarm_backend="device", rng="device", seed=3)
        opt.swarm_size = 4000
        np.random.seed(11)
        xs = [opt.optimize().copy() for _ in range(2)]
        sw = opt.swarms["expanders"]
        assert sw.use_graph == (mode == "1") and sw.in_kernel_exchange
        out[mode] = (xs, sw.best_positions.cpu().numpy().copy(), sw.best_values.cpu().numpy().copy(), opt.S.copy())
    for a, b in zip(out["1"][0], out["0"][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(out["1"][1], out["0"][1]) and np.array_equal(out["1"][2], out["0"][2]) and np.array_equal(out["1"][3], out["0"][3])


# ---------------------------------------------------------------- fp32 arithmetic mode (tcgen05 3xTF32, TMEM accumulators)
F32_TOL = 1e-4          # north star: "posterior mean/var within ... 1e-4 rel fp32" -- relative to the prior scale (sigma_f^2, max|y|)


@pytest.mark.parametrize("N,d,n,G", [(5, 1, 300, 1), (31, 2, 40, 1), (64, 2, 200, 1), (100, 2, 90, 2), (128, 2, 500, 3), (129, 2, 130, 1),
                                     (200, 3, 30, 1), (256, 4, 14, 2)])
def test_fp32_grid_kernel_matches_fp64_oracle(N, d, n, G):
    """so_posterior_grid_f32 against the fp64 oracle: |d mean| <= 1e-4 max(1, |y|), |d var| <= 1e-4 sigma_f^2, bounds within
    1e-4 * 2 sqrt(sigma_f^2); the safe bit is exactly the device's own l > fmin and equals the oracle's outside the band."""
    rs = np.random.RandomState(N + d)
    X = rs.uniform(-1.5, 1.5, (N, d))
    Y = np.stack([2 * np.exp(-np.sum(X * X, 1) / 8) + 0.05 * rs.randn(N) for _ in range(G)], axis=1)
    ls = rs.uniform(0.8, 1.4, d)
    var, noise, beta = 2.0, 0.05 ** 2, 2.0
    grid = sb.linearly_spaced_combinations([(-5, 5)] * d, n)
    M = grid.shape[0]
    from safeopt_b200.utilities import detect_grid
    eng = DeviceEngine(max_gps=G)
    eng.define_grid(detect_grid(grid))
    for i in range(G):
        eng.fit(i, X, Y[:, i], 0, ls, var, noise)
    eng.prepare_grid(0, 0, 0)
    eng.prepare_grid_f32(0, 0, M)
    Q, S = eng.empty((M, 2 * G)), eng.zeros((M,), "u8")
    mean, varr = eng.empty((G, M)), eng.empty((G, M))
    fmins = [0.1 * i for i in range(G)]
    eng.posterior_grid_f32(list(range(G)), 0, M, beta, fmins, means=[mean[i] for i in range(G)], variances=[varr[i] for i in range(G)],
                           Q=Q, q_cols=[2 * i for i in range(G)], S=S, safe_mode=_lib.SAFE_WRITE)
    Qh, Sh, mh, vh = Q.cpu().numpy(), S.cpu().numpy().astype(bool), mean.cpu().numpy(), varr.cpu().numpy()
    safe_o = np.ones(M, dtype=bool)
    clear = np.ones(M, dtype=bool)
    worst = [0.0, 0.0, 0.0]
    for i in range(G):
        go = gpy_lite.GPRegression(X, Y[:, [i]], kernel=gpy_lite.RBF(d, variance=var, lengthscale=ls, ARD=True), noise_var=noise)
        mo, vo = go.predict_noiseless(grid)
        mo, vo = mo[:, 0], vo[:, 0]
        lo, up = mo - beta * np.sqrt(vo), mo + beta * np.sqrt(vo)
        worst = [max(worst[0], np.abs(mh[i] - mo).max()), max(worst[1], np.abs(vh[i] - vo).max()), max(worst[2], np.abs(Qh[:, 2 * i] - lo).max())]
        assert np.abs(mh[i] - mo).max() <= F32_TOL * max(1.0, np.abs(Y).max())
        assert np.abs(vh[i] - vo).max() <= F32_TOL * var
        # near the variance floor sqrt amplifies: d sd = d var / (2 sd), so the bound on l/u is taken where sd is not tiny
        big = vo > 1e-3 * var
        assert np.abs(Qh[big, 2 * i] - lo[big]).max() <= F32_TOL * 2 * np.sqrt(var) * 3
        assert np.abs(Qh[big, 2 * i + 1] - up[big]).max() <= F32_TOL * 2 * np.sqrt(var) * 3
        assert np.array_equal(Qh[:, 2 * i], mh[i] - beta * np.sqrt(vh[i]))
        safe_o &= lo > fmins[i]
        clear &= np.abs(lo - fmins[i]) > F32_TOL * 2 * np.sqrt(var) * 3
    assert np.array_equal(Sh, np.all(Qh[:, ::2] > np.asarray(fmins), axis=1))
    assert np.array_equal(Sh[clear], safe_o[clear])
    print("fp32 N=%d d=%d M=%d G=%d: |d mean| %.2e |d var| %.2e |d l| %.2e, rows inside the band %d" % (N, d, M, G, *worst, int((~clear).sum())))
    # a shard of the rows gives the same bits as the whole grid
    if M > 300:
        r0, m = M // 3, M // 2
        eng.prepare_grid_f32(0, r0, m)
        Q2 = eng.empty((m, 2 * G))
        eng.posterior_grid_f32(list(range(G)), r0, m, beta, fmins, Q=Q2, q_cols=[2 * i for i in range(G)])
        assert torch.equal(Q2, Q[r0:r0 + m])
    eng.close()


def test_fp32_mode_of_safeopt_on_config3():
    """SafeOpt(precision='fp32') on BASELINE config 3 (2-D, 3 GPs, 500x500, N=128): bounds within the fp32 tolerance of the
    fp64 run, masks and query row identical outside the tolerance band, same class surface; a BO step keeps working (N = 129)."""
    w = workloads.config("C3")
    grid = sb.linearly_spaced_combinations(w.bounds, w.num_samples)
    out = {}
    for prec in ("fp64", "fp32"):
        gps = [sb.GPRegression(w.X, w.Y[:, [i]], kernel=sb.RBF(w.d, variance=w.variance, lengthscale=w.lengthscale, ARD=True), noise_var=w.noise_var)
               for i in range(3)]
        opt = sb.SafeOpt(gps, grid, w.fmin, beta=w.beta, threshold=w.threshold, precision=prec)
        x = opt.optimize()
        out[prec] = (opt.Q.copy(), opt.S.copy(), opt.M.copy(), opt.last_query_row, x.copy(), opt)
    Q64, S64, M64 = out["fp64"][:3]
    Q32, S32, M32 = out["fp32"][:3]
    tol = F32_TOL * 2 * np.sqrt(w.variance) * 3
    big = np.all((Q64[:, 1::2] - Q64[:, ::2]) > 4 * np.sqrt(1e-3 * w.variance), axis=1)
    assert np.abs(Q32 - Q64)[big].max() <= tol
    band = np.all(np.abs(Q64[:, ::2] - 0.0) > tol, axis=1)
    assert np.array_equal(S32[band], S64[band])
    best = Q64[S64, 0].max()
    band_m = band & (np.abs(Q64[:, 1] - best) > 2 * tol)
    assert np.array_equal(M32[band_m], M64[band_m])
    print("C3 fp32 vs fp64: |dQ| %.2e (tol %.1e), S differs on %d rows, M on %d rows (all inside the band), query rows %d / %d" % (
        np.abs(Q32 - Q64)[big].max(), tol, int((S32 != S64).sum()), int((M32 != M64).sum()), out["fp32"][3], out["fp64"][3]))
    opt = out["fp32"][5]
    assert any("fp32" in st for st in opt._grid_state.values())
    opt.add_new_data_point(out["fp32"][4], np.array([[1.0, 1.0, 1.0]]))
    assert opt.optimize() is not None and opt.t == 129


# ---------------------------------------------------------------- the boundary: foreign GPy-protocol models, the INTEGRATION.md stub
def test_foreign_gpy_protocol_models_go_straight_in():
    """"GPy model in, same methods out" (gp_opt.py:58-67, :347): objects that are NOT ours -- oracle.gpy_lite models with GPy's
    attribute surface, incl. a non-ARD kernel and a Prod of RBFs -- are handed to sb.SafeOpt / sb.SafeOptSwarm unchanged."""
    g = load_golden("expander_g2")
    gps, grid, fmin = golden_problem(g, "cpu")                      # gpy_lite.GPRegression objects
    assert not isinstance(gps[0], sb.GPRegression)
    opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
    x = opt.optimize()
    n_rows = int(g["n_rows"])
    assert opt.last_query_row == int(g["row_next"]) and np.array_equal(x, g["x_next"]) and np.abs(opt.Q - g["Q"]).max() < 1e-9 * 2 * np.sqrt(2.0)
    assert np.array_equal(opt.S, unpack_mask(g["S"], n_rows)) and np.array_equal(opt.M, unpack_mask(g["M"], n_rows))
    assert np.array_equal(opt.G, unpack_mask(g["G"], n_rows))
    opt.add_new_data_point(x, np.array([[1.0, 1.0]]))               # set_XY on the foreign objects, incremental device update
    assert gps[0].X.shape[0] == g["X"].shape[0] + 1 and opt.optimize() is not None and opt._fits.appends == len(gps)
    # contexts with a foreign Prod kernel
    c = load_golden("context_1p1c")
    kern = gpy_lite.RBF(1, variance=float(c["var0"]), lengthscale=float(c["ls0"]), active_dims=[0]) * \
        gpy_lite.RBF(1, variance=float(c["var1"]), lengthscale=float(c["ls1"]), active_dims=[1])
    assert type(kern).__name__ == "Prod"
    gp = gpy_lite.GPRegression(c["X"], c["Y"], kernel=kern, noise_var=float(c["noise_var"]))
    opt = sb.SafeOpt(gp, c["pset"], float(c["fmin"]), num_contexts=1, beta=float(c["beta"]), threshold=float(c["threshold"]))
    ctx = np.array([float(c["ctx0"])])
    xq = opt.optimize(context=ctx)
    n = int(c["n_rows"])
    assert np.abs(opt.Q - c["Q0"]).max() < 1e-9 * 2 * np.sqrt(3.0) and np.array_equal(xq, c["x0"])
    assert np.array_equal(opt.S, unpack_mask(c["S0"], n)) and np.array_equal(opt.M, unpack_mask(c["M0"], n))
    # a non-ARD kernel (one lengthscale for all dimensions) through the swarm optimiser
    rs = np.random.RandomState(2)
    X = rs.uniform(-0.5, 0.5, (12, 3))
    Y = (1 - 0.3 * np.sum(X * X, 1))[:, None]
    fgp = gpy_lite.GPRegression(X, Y, kernel=gpy_lite.RBF(3, variance=1.5, lengthscale=0.8), noise_var=1e-3)
    sw = sb.SafeOptSwarm(fgp, 0.0, bounds=[(-1, 1)] * 3, swarm_size=40)
    pts = rs.uniform(-1, 1, (50, 3))
    vd, sd = sw._compute_particle_fitness("expanders", pts)
    vo, so = port.particle_fitness([fgp], np.array([0.0]), 2.0, sw.scaling, "expanders", pts)
    assert np.abs(vd - vo).max() < 1e-9 and np.array_equal(sd, so)


def test_integration_md_stub_runs_against_the_port():
    """The ctypes stub INTEGRATION.md section 2 shows a maintainer (class B200Posterior) is executed verbatim against the
    port's GridProblem (the reference's SafeOpt attribute surface): same Q, same S as the port's own NumPy path."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"## 2\..*?```python\n(.*?)```", text, re.S).group(1)
    assert "class B200Posterior" in block
    block = block.replace('C.CDLL("libsafeopt_b200.so")', 'C.CDLL(%r)' % _lib.LIB_PATH)
    ns = {}
    exec(compile(block, "INTEGRATION.md#stub", "exec"), ns)
    g = load_golden("expander_g2")
    gps, grid, fmin = golden_problem(g, "cpu")
    prob = port.GridProblem.create(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]))
    prob.Q = np.zeros((grid.shape[0], 2 * len(gps)))
    prob.S = np.zeros(grid.shape[0], dtype=bool)
    ns["B200Posterior"](prob, device=torch.cuda.current_device()).update(prob, float(g["beta"]))
    assert np.abs(prob.Q - g["Q"]).max() < 1e-9 * 2 * np.sqrt(2.0)
    assert np.array_equal(prob.S, unpack_mask(g["S"], int(g["n_rows"])))


def test_scaled_operand_table_fallback_is_logged_and_equal(monkeypatch, capfd):
    """Above SO_APRIME_LIMIT_MB the grid path drops from the scaled-operand (A') kernel to the per-axis table kernel: it
    says so once on stderr and the answers do not change."""
    g = load_golden("config_C2")
    res = {}
    for limit in (None, "0"):
        if limit is None:
            monkeypatch.delenv("SO_APRIME_LIMIT_MB", raising=False)
        else:
            monkeypatch.setenv("SO_APRIME_LIMIT_MB", limit)
        gps, grid, fmin = golden_problem(g, "gpu")
        opt = sb.SafeOpt(gps[0], grid, fmin[0], beta=float(g["beta"]), threshold=float(g["threshold"]))
        opt.optimize()
        res[limit] = (opt.Q.copy(), opt.S.copy(), opt.M.copy(), opt.last_query_row)
    err = capfd.readouterr().err
    assert "scaled-operand table" in err
    a, b = res[None], res["0"]
    assert np.abs(a[0] - b[0]).max() < 1e-10 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]


# ---------------------------------------------------------------- multi-GPU (runs when the box has >= 2 GPUs)
def _nccl_worker(rank, world, port_no, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = load_golden("expander_g2")
        gps, grid, fmin = golden_problem(g, "gpu")
        opt = sb.SafeOpt(gps, grid, fmin, beta=float(g["beta"]), threshold=float(g["threshold"]), device=torch.device("cuda", rank))
        opt.optimize()
        n_rows = int(g["n_rows"])
        ok = (opt.last_query_row == int(g["row_next"]) and np.array_equal(opt.S, unpack_mask(g["S"], n_rows))
              and np.array_equal(opt.M, unpack_mask(g["M"], n_rows)) and np.array_equal(opt.G, unpack_mask(g["G"], n_rows))
              and np.abs(opt.Q - g["Q"]).max() < 1e-8)
        # more ranks than rows: the rank without rows still takes part in the in-kernel exchange (empty records)
        gp1 = sb.GPRegression(np.array([[0.0]]), np.array([[1.0]]), noise_var=0.01 ** 2)
        tiny = sb.SafeOpt(gp1, np.array([[0.05]]), fmin=0.0, device=torch.device("cuda", rank))
        x = tiny.optimize()
        ok = ok and tiny._row1 - tiny._row0 == (1 if rank == 0 else 0) and np.array_equal(x, [0.05]) and tiny._safe_info["n_safe"] == 1
        open(os.path.join(out, "ok_%d" % rank), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def _nccl_swarm_worker(rank, world, port_no, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ok = True
        for name in ["swarm_query_2d", "swarm_query_2d_mat32"]:
            g = load_golden(name)
            opt = _swarm_query_problem(g, swarm_backend="device", rng="host", device=torch.device("cuda", rank))
            np.random.seed(int(g["seed"]))
            x = opt.optimize()
            sw = opt.swarms["expanders"]
            ok = ok and sw.comm.world == world and sw.p1 - sw.p0 < int(g["swarm_size"])
            ok = ok and np.abs(x - g["x_next"]).max() < 1e-7 and opt.S.shape == g["S_final"].shape
            ok = ok and np.abs(opt.S - g["S_final"]).max() < 1e-7
            full = D.gather_padded_rows(sw.comm, sw.best_positions, int(g["swarm_size"])).cpu().numpy()
            ok = ok and np.abs(full - g["expanders_best_positions"]).max() < 1e-7
        open(os.path.join(out, "ok_%d" % rank), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def _nccl_device_rng_worker(rank, world, port_no, out):
    """rng='device' draws by global particle index: the sharded swarm must follow the single-GPU trajectory bit for bit
    (graph replay + in-kernel record exchange on both sides)."""
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = load_golden("swarm_query_2d")
        res = []
        for distributed in (True, False):
            opt = _swarm_query_problem(g, swarm_backend="device", rng="device", seed=5, device=torch.device("cuda", rank),
                                       distributed=distributed)
            opt.swarm_size = 3001
            np.random.seed(4)
            x = opt.optimize()
            sw = opt.swarms["expanders"]
            full = D.gather_padded_rows(sw.comm, sw.best_positions, 3001).cpu().numpy()
            res.append((x.copy(), full, opt.S.copy(), sw.comm.world, sw.in_kernel_exchange))
        ok = res[0][3] == world and res[1][3] == 1 and res[0][4] and res[1][4]
        ok = ok and np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
        open(os.path.join(out, "ok_%d" % rank), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def test_graph_capture_survives_garbage_of_earlier_optimisers(monkeypatch):
    """SafeOptSwarm <-> DeviceSwarm is a reference cycle, so an abandoned optimiser (and its engine: cudaFree / cudaFreeHost in
    so_destroy) dies only in Python's cyclic collector.  If that collection ran while a PSO iteration is being captured into a
    CUDA graph, the capture would be invalidated ('operation not permitted when stream is capturing' -> torch.AcceleratorError):
    the capture therefore runs with the collector switched off.  Here two optimisers become garbage INSIDE the capture and an
    automatic collection is simulated at that point (collect iff the collector is enabled)."""
    import gc
    from safeopt_b200.swarm import DeviceSwarm
    g = load_golden("swarm_query_2d")
    victims = []
    for _ in range(2):
        o = _swarm_query_problem(g, swarm_backend="device", rng="device", seed=3)
        o.optimize()
        victims.append(o)
    torch.cuda.synchronize()
    orig = DeviceSwarm._iteration_dev
    seen = {"captures": 0, "collections": 0}

    def patched(self):
        if torch.cuda.is_current_stream_capturing():
            seen["captures"] += 1
            if victims:
                victims.pop()
            if gc.isenabled():
                seen["collections"] += 1
                gc.collect()
        orig(self)

    monkeypatch.setattr(DeviceSwarm, "_iteration_dev", patched)
    assert gc.isenabled()
    test_device_swarm_device_rng_runs()
    torch.cuda.synchronize()
    assert gc.isenabled()                                   # switched back on after every capture
    assert seen["captures"] >= 3 and seen["collections"] == 0 and not victims
    gc.collect()                                            # the abandoned engines are finalised here, outside any capture
    torch.cuda.synchronize()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_device_rng_swarm_equals_single_gpu(tmp_path):
    import os
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_device_rng_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    assert all(open(os.path.join(str(tmp_path), "ok_%d" % r)).read() == "1" for r in range(2))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_swarm_matches_golden(tmp_path):
    import os
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_swarm_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    assert all(open(os.path.join(str(tmp_path), "ok_%d" % r)).read() == "1" for r in range(2))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_optimize_matches_golden(tmp_path):
    import os
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    assert all(open(os.path.join(str(tmp_path), "ok_%d" % r)).read() == "1" for r in range(2))
