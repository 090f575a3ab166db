"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, grid
recognition, hyper-parameter extraction, the reference's own API-surface tests (which need no GP
arithmetic), and loud failure without a GPU."""
import os
import re

import numpy as np
import pytest

import safeopt_b200 as sb
from safeopt_b200 import _lib, gpmodel
from safeopt_b200.distributed import combine_max_first, shard_bounds
from safeopt_b200.gp_opt import GaussianProcessOptimization
from safeopt_b200.utilities import detect_grid, grid_row_strides, grid_rows_from_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "safeopt_b200.h")).read()
    declared = set(re.findall(r"\b(so_[a-z_0-9]+)\s*\(", header))
    declared -= {"so_handle"}
    assert len(declared) >= 20
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libsafeopt_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "ctypes binding lacks %s" % name
    assert set(_lib.SIGNATURES) == declared
    assert lib.so_abi_version() == _lib.ABI_VERSION
    assert _lib.status_string(_lib.SO_ERR_NOT_PD) == "covariance matrix not positive definite"


@pytest.mark.parametrize("NB", [1, 7, 32, 33, 35, 36, 40, 52, 63, 64, 65, 100, 129, 256])
def test_block_row_plan_is_a_balanced_partition(NB):
    """plan_rows (host side of the contraction kernels): every block row is owned by exactly one warp, rows ascend inside a
    pass, unused slots trail, and the triangular work (row i costs i + 1 k-blocks) is balanced over the eight warps."""
    import ctypes
    lib = _lib.load()
    table = np.full(8 * 8 * 4, -2, dtype=np.int16)
    npass = ctypes.c_int(0)
    assert lib.so_debug_row_plan(NB, table.ctypes.data, ctypes.byref(npass)) == 0
    t = table.reshape(8, 8, 4)
    n = npass.value
    assert n == -(-(-(-NB // 8)) // 4) and np.all(t[n:] == -1)
    rows = t[:n][t[:n] >= 0]
    assert sorted(rows.tolist()) == list(range(NB))
    for p in range(n):
        for w in range(8):
            slot = t[p, w]
            used = slot[slot >= 0]
            assert np.all(np.diff(used) > 0) and np.all(slot[len(used):] == -1)
    load = np.array([sum(int(r) + 1 for r in t[:n, w].ravel() if r >= 0) for w in range(8)])
    ideal = NB * (NB + 1) / 2 / 8
    assert load.max() <= ideal + NB          # longest-first greedy: within one (longest) row of the ideal share
    assert lib.so_debug_row_plan(0, table.ctypes.data, ctypes.byref(npass)) == _lib.SO_ERR_BAD_ARG


@pytest.mark.parametrize("NB", [33, 36, 37, 40, 44, 47, 48, 49, 96])
def test_six_row_plan_is_a_balanced_partition(NB):
    """The same plan with six block rows per warp and pass (grid kernel, NB = 36..48: the whole triangle in ONE pass)."""
    import ctypes
    lib = _lib.load()
    table = np.full(8 * 8 * 6, -2, dtype=np.int16)
    npass = ctypes.c_int(0)
    assert lib.so_debug_row_plan_slots(NB, 6, table.ctypes.data, ctypes.byref(npass)) == 0
    t = table.reshape(8, 8, 6)
    n = npass.value
    assert n == -(-(-(-NB // 8)) // 6) and np.all(t[n:] == -1) and (n == 1) == (NB <= 48)
    rows = t[:n][t[:n] >= 0]
    assert sorted(rows.tolist()) == list(range(NB))
    for p in range(n):
        for w in range(8):
            slot = t[p, w]
            used = slot[slot >= 0]
            assert np.all(np.diff(used) > 0) and np.all(slot[len(used):] == -1)
    load = np.array([sum(int(r) + 1 for r in t[:n, w].ravel() if r >= 0) for w in range(8)])
    assert load.max() <= NB * (NB + 1) / 2 / 8 + NB
    # four slots through the same entry point give the plan of so_debug_row_plan
    t4 = np.full(8 * 8 * 4, -2, dtype=np.int16)
    n4 = ctypes.c_int(0)
    assert lib.so_debug_row_plan(NB, t4.ctypes.data, ctypes.byref(n4)) == 0
    assert lib.so_debug_row_plan_slots(NB, 4, table.ctypes.data, ctypes.byref(npass)) == 0 and npass.value == n4.value
    assert np.array_equal(table.reshape(8, 8, 6)[:, :, :4], t4.reshape(8, 8, 4)) and np.all(table.reshape(8, 8, 6)[:, :, 4:] == -1)
    assert lib.so_debug_row_plan_slots(NB, 5, table.ctypes.data, ctypes.byref(npass)) == _lib.SO_ERR_BAD_ARG


def test_tile_plans_fit_the_device_for_every_size():
    """Host-side planning of the posterior kernels for every NB (N = 8 .. 2048) and 0..3 further outputs, on a B200's
    227 KB of opt-in shared memory: whatever is planned fits, tiles are whole 8-row blocks, the eight warps are split
    RG x CG, the passes cover all block rows, and the grid kernel keeps 48-row tiles wherever the ring is used."""
    lib = _lib.load()
    limit, sms = 232448, 148
    out = np.zeros(18, dtype=np.int64)
    seen_ring = seen_db48 = 0
    for nb in range(1, 257):
        for n_extra in range(4):
            assert lib.so_debug_tile_plans(nb, 4, 6_250_000, n_extra, limit, sms, out.ctypes.data) == 0
            st, bt, rg, cg, T, npass, ring, stages, smem, warps = out[:10]
            if st == 0:
                # the tile is planned for one output; with further outputs a launch that no longer fits is refused with
                # SO_ERR_CAPACITY and the host evaluates the GPs one by one (the ring re-sizes itself and always fits)
                assert (smem <= limit or (n_extra > 0 and not ring)) and bt in (2, 4, 6) and rg * cg == warps == 8 and T == 8 * bt * cg
                assert out[17] in (2, 4, 6) and out[17] * rg * npass >= nb
                if ring:
                    assert stages >= 4 and T == 48 and rg == 8
                    seen_ring += 1
                elif T == 48 and cg == 1:
                    seen_db48 += 1
            else:
                assert st == _lib.SO_ERR_CAPACITY
            st2, bt2, rg2, cg2, T2, npass2, smem2 = out[10:17]
            if st2 == 0:
                assert smem2 <= limit and bt2 in (2, 4, 8) and rg2 * cg2 == 8 and T2 == 8 * bt2 * cg2 and 4 * rg2 * npass2 >= nb
            else:
                assert st2 == _lib.SO_ERR_CAPACITY and nb > 150           # the resident K tile caps the explicit-rows kernel
    assert seen_ring > 0 and seen_db48 > 0
    # config 4: N = 256 -> 48-row double buffer; N = 257..280 keep it; N = 281..384: 32-row tile, six block rows per warp in
    # one pass; N = 512 streams
    for nb, want_T, want_ring, want_ns, want_pass in [(32, 48, 0, 4, 1), (33, 48, 0, 4, 2), (35, 48, 0, 4, 2), (36, 32, 0, 6, 1),
                                                      (48, 32, 0, 6, 1), (49, 32, 0, 4, 2), (52, 32, 0, 4, 2), (64, 48, 1, 4, 2)]:
        lib.so_debug_tile_plans(nb, 4, 6_250_000, 0, limit, sms, out.ctypes.data)
        assert (out[4], out[6], out[17], out[5]) == (want_T, want_ring, want_ns, want_pass), nb


def test_record_layouts_match_header():
    import ctypes
    from safeopt_b200.engine import MAX_REC_DTYPE, SAFE_REC_DTYPE
    assert ctypes.sizeof(_lib.SafeRecord) == SAFE_REC_DTYPE.itemsize == 64
    assert ctypes.sizeof(_lib.MaxRecord) == MAX_REC_DTYPE.itemsize == 64
    assert [f[0] for f in _lib.SafeRecord._fields_] == list(SAFE_REC_DTYPE.names)
    assert [f[0] for f in _lib.MaxRecord._fields_] == list(MAX_REC_DTYPE.names)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    gp = sb.GPRegression(np.zeros((1, 1)), np.ones((1, 1)))
    grid = sb.linearly_spaced_combinations([(-1, 1)], 10)
    with pytest.raises(sb.NativeLibraryError):
        sb.SafeOpt(gp, grid, 0.0)
    with pytest.raises(sb.NativeLibraryError):
        gp.predict_noiseless(grid)
    with pytest.raises(sb.NativeLibraryError):
        sb.SafeOptSwarm(gp, 0.0, bounds=[(-1, 1)])


class _RecordingEngine:
    """Stands in for DeviceEngine in the host-logic test of _DeviceFits: records which fit entry point would be called."""

    def __init__(self, append_ok=True):
        self.calls = []
        self.append_ok = append_ok

    def fit(self, i, X, Y, kind, ls, variance, noise):
        self.calls.append(("fit", i, X.shape[0]))

    def fit_append(self, i, x, y):
        self.calls.append(("append", i, float(y)))
        return self.append_ok

    def fit_remove_last(self, i):
        self.calls.append(("remove", i))


def test_device_fits_choose_the_cheapest_update(monkeypatch):
    """_DeviceFits.refresh: nothing changed -> no call; one row appended -> so_fit_append (falling back to so_fit when the
    device refuses); last row removed -> so_fit_remove_last; anything else -> so_fit.  GPs that share inputs, kernel and
    noise form one group (one launch), a GP with its own kernel its own."""
    from safeopt_b200.gp_opt import _DeviceFits
    monkeypatch.delenv("SAFEOPT_B200_INCREMENTAL_FIT", raising=False)
    monkeypatch.delenv("SAFEOPT_B200_SHARE_FITS", raising=False)
    rs = np.random.RandomState(0)
    X, Y = rs.rand(5, 2), rs.rand(5, 3)
    mk = lambda i, var=2.0: sb.GPRegression(X, Y[:, [i]], kernel=sb.RBF(2, variance=var, lengthscale=np.ones(2), ARD=True), noise_var=0.01)
    gps = [mk(0), mk(1), mk(2, var=1.5)]
    eng = _RecordingEngine()
    fits = _DeviceFits(eng, gps)
    seen = []
    fits.refresh(lambda i, hyper: seen.append(i))
    assert eng.calls == [("fit", 0, 5), ("fit", 1, 5), ("fit", 2, 5)] and seen == [0, 1, 2]
    assert fits.groups == [[0, 1], [2]]
    eng.calls.clear()
    fits.refresh()
    assert eng.calls == []
    x_new, y_new = rs.rand(1, 2), rs.rand(1, 3)
    for i, gp in enumerate(gps):
        gp.set_XY(np.vstack([gp.X, x_new]), np.vstack([gp.Y, y_new[:, [i]]]))
    fits.refresh()
    assert eng.calls == [("append", i, float(y_new[0, i])) for i in range(3)] and fits.appends == 3
    assert fits.groups == [[0, 1], [2]]
    eng.calls.clear()
    gps[0].set_XY(gps[0].X[:-1], gps[0].Y[:-1])
    fits.refresh()
    assert eng.calls == [("remove", 0)] and fits.groups == [[0], [1], [2]]          # GP 0 now has different data
    eng.calls.clear()
    gps[1].kern.lengthscale = np.array([0.5, 2.0])                                  # hyper-parameter change
    Xm = gps[2].X.copy()
    Xm[0, 0] += 1e-9                                                                # a row in the middle changed
    gps[2].set_XY(Xm, gps[2].Y)
    fits.refresh()
    assert eng.calls == [("fit", 1, 6), ("fit", 2, 6)]
    eng.calls.clear()
    eng.append_ok = False                                                           # device buffers full
    gps[1].set_XY(np.vstack([gps[1].X, x_new]), np.vstack([gps[1].Y, y_new[:, [1]]]))
    fits.refresh()
    assert eng.calls == [("append", 1, float(y_new[0, 1])), ("fit", 1, 7)]
    monkeypatch.setenv("SAFEOPT_B200_SHARE_FITS", "0")
    assert _DeviceFits(_RecordingEngine(), [mk(0), mk(1)]).share is False


def test_weak_scaling_workload_shapes():
    """bench.py --scaling weak multiplies the points of axis 1 (the slowest axis of the reference row order) by the rank
    count, so that contiguous row blocks of equal size are whole axis-1 slabs."""
    from safeopt_b200 import workloads
    from safeopt_b200.distributed import shard_bounds
    from safeopt_b200.utilities import grid_row_strides
    w = workloads.config("C4")
    assert w.samples_per_axis == [50] * 4 and w.n_rows == 50 ** 4
    per = w.samples_per_axis
    per[1] *= 8
    w.num_samples = per
    assert w.n_rows == 8 * 50 ** 4
    strides = grid_row_strides(w.samples_per_axis)
    assert strides[1] == max(strides) == 50 ** 3
    for r in range(8):
        lo, hi = shard_bounds(w.n_rows, 8, r)
        assert hi - lo == 50 ** 4 and lo % (50 * strides[1]) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; no module of the product may import it."""
    pkg = os.path.join(ROOT, "safeopt_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


# ---------------------------------------------------------------- grid recognition
@pytest.mark.parametrize("bounds,n", [([(-1, 1)], 7), ([(-5, 5)] * 2, [4, 5]), ([(-5, 5), (0, 1), (2, 3)], [3, 4, 5]),
                                      ([(-5, 5)] * 4, [3, 2, 4, 5]), ([(0, 1)] * 5, 3)])
def test_detect_grid_roundtrip(bounds, n):
    from oracle import safeopt_port as port
    grid = sb.linearly_spaced_combinations(bounds, n)
    assert np.array_equal(grid, port.linearly_spaced_combinations(bounds, n))
    axes = detect_grid(grid)
    assert axes is not None
    ns = n if isinstance(n, list) else [n] * len(bounds)
    assert [len(a) for a in axes] == ns
    for a, (lo, hi), k in zip(axes, bounds, ns):
        assert np.array_equal(a, np.linspace(lo, hi, k))
    assert np.array_equal(grid_rows_from_index(axes, np.arange(grid.shape[0])), grid)
    strides = grid_row_strides(ns)
    if len(ns) >= 2:
        assert strides[1] == max(strides)


def test_detect_grid_rejects_non_grids():
    grid = sb.linearly_spaced_combinations([(-5, 5)] * 3, 4).copy()
    bad = grid.copy()
    bad[5, 1] += 1e-9
    assert detect_grid(bad) is None
    assert detect_grid(np.random.RandomState(0).rand(64, 3)) is None
    assert detect_grid(grid[:-1]) is None
    assert detect_grid(grid.astype(np.float32)) is None
    assert detect_grid(sb.linearly_spaced_combinations([(0, 1)] * 7, 2)) is None      # beyond the grid fast path (d <= 6)


def test_linearly_spaced_combinations_api():
    g = sb.linearly_spaced_combinations([(-1, 1), (0, 2)], 3)
    assert g.shape == (9, 2)
    g = sb.linearly_spaced_combinations([(-1, 1), (0, 2)], [2, 5])
    assert g.shape == (10, 2)


# ---------------------------------------------------------------- sharding / record combine
def test_shard_bounds_cover_rows_exactly():
    for M in [0, 1, 7, 100, 6_250_000]:
        for world in [1, 2, 3, 8]:
            spans = [shard_bounds(M, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1


def test_combine_max_first_is_numpy_argmax():
    rs = np.random.RandomState(3)
    for _ in range(50):
        v = rs.randint(0, 4, size=8).astype(float)
        rows = rs.permutation(100)[:8]
        val, row = combine_max_first(v, rows)
        assert val == v.max() and row == rows[v == v.max()].min()
    assert combine_max_first([1.0, 5.0], [-1, -1]) == (-np.inf, -1)
    assert combine_max_first([1.0, -np.inf], [4, -1]) == (1.0, 4)


def test_candidate_order_on_ties_follows_the_reference():
    """gp_opt.py:545-552 visits ``keys.argsort()[::-1]``: on exactly equal keys (where NumPy's sort is stable, i.e. the
    insertion sort it runs up to 16 elements) the HIGHER row comes first."""
    from safeopt_b200.distributed import Comm, order_candidates
    keys = np.array([1.0, 2.0, 2.0, 1.0, 3.0, 2.0])
    rows = np.array([10, 11, 12, 13, 14, 15])
    ref = rows[keys.argsort()[::-1]]
    assert list(ref) == [14, 15, 12, 11, 13, 10]
    # single rank: the caller pre-sorts (rows descending, then stable by key descending) exactly like SafeOpt._ordered_candidates
    import torch
    r, k = torch.from_numpy(rows), torch.from_numpy(keys)
    o = torch.argsort(r, descending=True)
    r, k = r[o], k[o]
    o = torch.argsort(k, descending=True, stable=True)
    assert np.array_equal(order_candidates(Comm(), r[o].numpy(), k[o].numpy()), ref)


def test_fmin_none_means_unconstrained():
    """The docstring promises None == -inf; np.asarray(None, float) would be NaN and make every row unsafe."""
    from safeopt_b200.gp_opt import GaussianProcessOptimization

    class K:
        def Kdiag(self, x):
            return np.ones(len(x))

    class GP:
        X, Y, input_dim, kern = np.zeros((1, 1)), np.zeros((1, 1)), 1, K()

    opt = GaussianProcessOptimization([GP(), GP()], fmin=[None, 0.5])
    assert np.array_equal(opt.fmin, [-np.inf, 0.5])
    assert np.array_equal(GaussianProcessOptimization(GP(), fmin=None).fmin, [-np.inf])
    with pytest.raises(ValueError):
        GaussianProcessOptimization(GP(), fmin=float("nan"))


# ---------------------------------------------------------------- hyper-parameter adapter
def test_extract_hyper_from_product_and_oracle_models():
    from oracle import gpy_lite
    X = np.random.RandomState(0).rand(5, 3)
    Y = np.zeros((5, 1))
    gp = gpy_lite.GPRegression(X, Y, kernel=gpy_lite.Matern52(3, variance=1.5, lengthscale=[1, 2, 3], ARD=True), noise_var=0.1)
    h = gpmodel.extract_hyper(gp)
    assert h.kind == _lib.KERNEL_MATERN52 and np.array_equal(h.lengthscale, [1, 2, 3]) and h.variance == 1.5 and h.noise_var == 0.1
    gp = gpy_lite.GPRegression(X, Y, kernel=gpy_lite.RBF(3, variance=2.0, lengthscale=0.7), noise_var=0.2)
    h = gpmodel.extract_hyper(gp)
    assert h.kind == _lib.KERNEL_RBF and np.allclose(h.lengthscale, 0.7) and h.lengthscale.shape == (3,)
    # context example shape: k_param (dims 0,1) * k_context (dim 2)
    prod = gpy_lite.RBF(2, variance=2.0, lengthscale=[1.0, 0.5], ARD=True, active_dims=[0, 1]) * \
        gpy_lite.RBF(1, variance=3.0, lengthscale=4.0, active_dims=[2])
    gp = gpy_lite.GPRegression(X, Y, kernel=prod, noise_var=0.2)
    h = gpmodel.extract_hyper(gp)
    assert h.kind == _lib.KERNEL_RBF and np.array_equal(h.lengthscale, [1.0, 0.5, 4.0]) and h.variance == 6.0
    bad = gpy_lite.RBF(2, active_dims=[0, 1]) * gpy_lite.Matern32(1, active_dims=[2])
    with pytest.raises(gpmodel.UnsupportedModelError):
        gpmodel.extract_hyper(gpy_lite.GPRegression(X, Y, kernel=bad))

    class Weird:
        input_dim = 3
        lengthscale = np.ones(1)
        variance = np.ones(1)
    gp.kern = Weird()
    with pytest.raises(gpmodel.UnsupportedModelError):
        gpmodel.extract_hyper(gp)


def test_extract_hyper_refuses_sums_mean_functions_and_normalizers():
    """A GPy Add kernel (name 'sum') also has ``.parts``; folding it into one ARD RBF would silently change the posterior
    of a safety-critical optimiser.  Same for models whose predict applies a mean function or an output normaliser."""
    from oracle import gpy_lite
    X = np.random.RandomState(0).rand(5, 2)
    Y = np.zeros((5, 1))

    class Add:                      # test double with GPy's Add surface
        name = "sum"
        input_dim = 2

        def __init__(self, parts):
            self.parts = parts

    gp = gpy_lite.GPRegression(X, Y, kernel=gpy_lite.RBF(2), noise_var=0.1)
    gp.kern = Add([gpy_lite.RBF(1, variance=1.0, active_dims=[0]), gpy_lite.RBF(1, variance=2.0, active_dims=[1])])
    with pytest.raises(gpmodel.UnsupportedModelError, match="composite"):
        gpmodel.extract_hyper(gp)
    gp = gpy_lite.GPRegression(X, Y, kernel=gpy_lite.RBF(2), noise_var=0.1)
    assert gpmodel.extract_hyper(gp).kind == _lib.KERNEL_RBF
    gp.mean_function = lambda x: x
    with pytest.raises(gpmodel.UnsupportedModelError, match="mean_function"):
        gpmodel.extract_hyper(gp)
    gp.mean_function = None
    gp.normalizer = object()
    with pytest.raises(gpmodel.UnsupportedModelError, match="normalizer"):
        gpmodel.extract_hyper(gp)
    gp.normalizer = False           # GPy's "off"
    assert gpmodel.extract_hyper(gp).variance == 1.0


def test_host_kernels_match_oracle():
    from oracle import gpy_lite
    X = np.random.RandomState(1).randn(6, 2)
    Z = np.random.RandomState(2).randn(4, 2)
    for a, b in [(sb.RBF, gpy_lite.RBF), (sb.Matern32, gpy_lite.Matern32), (sb.Matern52, gpy_lite.Matern52)]:
        ka = a(2, variance=1.7, lengthscale=[0.6, 1.4], ARD=True)
        kb = b(2, variance=1.7, lengthscale=[0.6, 1.4], ARD=True)
        assert np.allclose(ka.K(X, Z), kb.K(X, Z), atol=1e-13)
        assert np.allclose(ka.K(X), kb.K(X), atol=1e-13)
        assert np.allclose(ka.Kdiag(X), kb.Kdiag(X))


# ---------------------------------------------------------------- the reference's own unit tests, on our classes
class TestGPOptimizationSurface(object):
    """Mirrors /root/reference/safeopt/tests/test_gps.py (no GP arithmetic involved => runs on CPU)."""

    @pytest.fixture
    def gps(self):
        gp1 = sb.GPRegression(np.array([[0]]), np.array([[0]]), kernel=sb.RBF(1, variance=2))
        gp2 = sb.GPRegression(np.array([[0]]), np.array([[0]]), kernel=sb.Matern32(1, variance=4))
        return gp1, gp2

    def test_init(self, gps):
        gp1, _ = gps
        opt = GaussianProcessOptimization(gp1, fmin=0, beta=2, num_contexts=1, threshold=0, scaling="auto")
        assert opt.beta(0) == 2
        opt = GaussianProcessOptimization(gp1, fmin=[0], beta=lambda x: 5, num_contexts=1, threshold=0, scaling="auto")
        assert opt.beta(10) == 5

    def test_multi_init(self, gps):
        opt = GaussianProcessOptimization(list(gps), fmin=0, beta=2, num_contexts=1, threshold=0, scaling="auto")
        np.testing.assert_allclose(opt.scaling, np.array([np.sqrt(2), np.sqrt(4)]))

    def test_scaling(self, gps):
        pytest.raises(ValueError, GaussianProcessOptimization, list(gps), 2, scaling=[5])
        opt = GaussianProcessOptimization(list(gps), fmin=[1, 0], beta=2, num_contexts=1, threshold=0, scaling=[1, 2])
        np.testing.assert_allclose(opt.scaling, np.array([1, 2]))

    def test_data_adding(self, gps):
        gp1, gp2 = gps
        gp1.set_XY(np.array([[0.]]), np.array([[1.]]))
        opt = GaussianProcessOptimization(gp1, 0)
        opt.add_new_data_point(2, 3)
        x, y = opt.data
        np.testing.assert_allclose(x, np.array([[0], [2]]))
        np.testing.assert_allclose(y, np.array([[1], [3]]))
        gp1.set_XY(np.array([[0.]]), np.array([[1.]]))
        gp2.set_XY(np.array([[0.]]), np.array([[11.]]))
        opt = GaussianProcessOptimization([gp1, gp2], [0, 1])
        opt.add_new_data_point(2, [2, 3])
        x, y = opt.data
        np.testing.assert_allclose(x, np.array([[0], [2]]))
        np.testing.assert_allclose(y, np.array([[1, 11], [2, 3]]))
        opt.add_new_data_point(3, [2, np.nan])
        np.testing.assert_allclose(opt.x, np.array([[0], [2], [3]]))
        np.testing.assert_allclose(opt.y, np.array([[1, 11], [2, 3], [2, np.nan]]))
        for i, gp in enumerate(opt.gps):
            not_nan = ~np.isnan(opt.y[:, i])
            np.testing.assert_allclose(gp.X, opt.x[not_nan, :])
            np.testing.assert_allclose(gp.Y[:, 0], opt.y[not_nan, i])
        opt.remove_last_data_point()
        np.testing.assert_allclose(opt.x, np.array([[0], [2]]))
        np.testing.assert_allclose(opt.y, np.array([[1, 11], [2, 3]]))
        for i, gp in enumerate(opt.gps):
            not_nan = ~np.isnan(opt.y[:, i])
            np.testing.assert_allclose(gp.X, opt.x[not_nan, :])
            np.testing.assert_allclose(gp.Y[:, 0], opt.y[not_nan, i])

    def test_contexts(self):
        gp1 = sb.GPRegression(np.array([[0, 0]]), np.array([[5]]), kernel=sb.RBF(2, variance=2))
        gp2 = sb.GPRegression(np.array([[0, 0]]), np.array([[6]]), kernel=sb.Matern32(2, variance=4))
        opt = GaussianProcessOptimization([gp1, gp2], fmin=[0, 0], num_contexts=1)
        opt.add_new_data_point(1, [3, 4], context=2)
        np.testing.assert_allclose(opt.x, np.array([[0, 0], [1, 2]]))
        np.testing.assert_allclose(opt.y, np.array([[5, 6], [3, 4]]))
        for i, gp in enumerate(opt.gps):
            np.testing.assert_allclose(gp.X, opt.x)
            np.testing.assert_allclose(gp.Y[:, 0], opt.y[:, i])


def test_workloads_are_deterministic():
    from safeopt_b200 import workloads
    a, b = workloads.config("C4"), workloads.config("C4")
    assert np.array_equal(a.X, b.X) and np.array_equal(a.Y, b.Y) and a.n_rows == 50 ** 4 and a.n_train == 256
    c3 = workloads.config("C3")
    assert c3.Y.shape == (128, 3) and c3.n_rows == 250000
    sw = workloads.swarm_workload(1000, 64)
    assert sw.particles.shape == (1000, 6) and sw.Y.shape == (64, 2)
