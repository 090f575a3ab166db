"""Shared pytest configuration: the ``gpu`` marker and helpers to rebuild golden problems."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def unpack_mask(packed, n):
    return np.unpackbits(packed)[:n].astype(bool)


GRID_CASES = ["doctest_1d", "config_C1", "config_C2", "config_C3_n80", "config_C4_n10", "expander_g1", "expander_g2",
              "expander_tight", "full_sets_g1", "matern32_3d", "matern52_2d_g2", "lipschitz_g1", "lipschitz_g2"]


def golden_lipschitz(g):
    """Lipschitz constants of a grid fixture as the constructor argument (None if the GP rule is used)."""
    lip = np.atleast_1d(g["lipschitz"]) if "lipschitz" in g.files else np.zeros(0)
    if lip.size == 0:
        return None
    return [float(v) for v in lip] if lip.size > 1 else float(lip[0])


def oracle_kernel(kind, d, variance, ls):
    from oracle import gpy_lite
    cls = {0: gpy_lite.RBF, 1: gpy_lite.Matern32, 2: gpy_lite.Matern52}[int(kind)]
    return cls(d, variance=float(variance), lengthscale=np.asarray(ls, dtype=float), ARD=True)


def device_kernel(kind, d, variance, ls):
    import safeopt_b200 as sb
    cls = {0: sb.RBF, 1: sb.Matern32, 2: sb.Matern52}[int(kind)]
    return cls(d, variance=float(variance), lengthscale=np.asarray(ls, dtype=float), ARD=True)


def golden_problem(g, which):
    """Rebuild (gps, grid, fmin) of a grid fixture with oracle ('cpu') or device ('gpu') models."""
    from oracle import safeopt_port as port
    X, Y = g["X"], g["Y"]
    d = X.shape[1]
    bounds = [tuple(b) for b in g["bounds"]]
    n = [int(v) for v in np.atleast_1d(g["num_samples"])]
    grid = port.linearly_spaced_combinations(bounds, n)
    if which == "cpu":
        from oracle import gpy_lite
        gps = [gpy_lite.GPRegression(X, Y[:, [i]], kernel=oracle_kernel(g["kind"], d, g["variance"], g["lengthscale"]),
                                     noise_var=float(g["noise_var"])) for i in range(Y.shape[1])]
    else:
        import safeopt_b200 as sb
        gps = [sb.GPRegression(X, Y[:, [i]], kernel=device_kernel(g["kind"], d, g["variance"], g["lengthscale"]),
                               noise_var=float(g["noise_var"])) for i in range(Y.shape[1])]
    return gps, grid, [float(v) for v in g["fmin"]]


@pytest.fixture(scope="session")
def reference_pkg():
    from oracle import compat
    if not compat.reference_available():
        pytest.skip("reference tree not present (only in the build container)")
    import warnings
    warnings.simplefilter("ignore")
    return compat.import_reference()
