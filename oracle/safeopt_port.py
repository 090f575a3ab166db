"""NumPy restatement of the SafeOpt hot path.  TEST INFRASTRUCTURE ONLY (oracle).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this.  It exists because ``/root/reference`` cannot travel
to the GPU box; every function cites the reference lines it restates and
``tests/test_oracle.py`` checks it, in the build container, against the reference's own
``safeopt`` package imported through :mod:`oracle.compat` (masks and query points equal,
``Q`` bit-identical) and against the committed golden fixtures everywhere.

The GP arithmetic underneath is :mod:`oracle.gpy_lite` (GPy restatement; "parity unpinned"
at that level -- see its header).

Row order, comparison operators (strict ``>`` for S, ``>=`` for M and the expander test),
the unscaled sort key, first-index argmax tie-breaks and the refit-based expander test all
follow the reference, not the GPU implementation (which uses L^-1 and a rank-1 update).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np
from scipy.special import expit
from scipy.stats import norm


# --------------------------------------------------------------------------- grid (a1)
def linearly_spaced_combinations(bounds, num_samples):
    """Cartesian grid in the reference's row order (safeopt/utilities.py:21-54).

    d == 1: ``linspace`` as a column.  d >= 2: ``meshgrid`` with its default 'xy'
    indexing, ravelled and transposed -- hence axis 1 is the slowest varying, then
    axis 0, then axes 2..d-1 (fastest)."""
    d = len(bounds)
    if np.isscalar(num_samples):
        num_samples = [int(num_samples)] * d
    axes = [np.linspace(lo, hi, int(n)) for (lo, hi), n in zip(bounds, num_samples)]
    if d == 1:
        return axes[0][:, None]
    mesh = np.meshgrid(*axes)
    return np.array([m.ravel() for m in mesh]).T


# --------------------------------------------------------------------------- CI + S (a4, a7)
def grid_rows(bounds, num_samples, rows):
    """Rows ``rows`` of ``linearly_spaced_combinations(bounds, num_samples)`` without building the whole grid
    (bit-identical: meshgrid only replicates the linspace values; row order of utilities.py:50-54:
    variable 1 slowest, then variable 0, then variables 2..d-1)."""
    d = len(bounds)
    if not isinstance(num_samples, Sequence):
        num_samples = [num_samples] * d
    axes = [np.linspace(b[0], b[1], int(n)) for b, n in zip(bounds, num_samples)]
    rows = np.asarray(rows, dtype=np.int64)
    order = [0] if d == 1 else [1, 0] + list(range(2, d))       # slowest ... fastest
    out = np.empty((rows.size, d))
    stride = 1
    for j in reversed(order):
        out[:, j] = axes[j][(rows // stride) % len(axes[j])]
        stride *= len(axes[j])
    return out


def confidence_intervals(gps, inputs, beta, chunk: Optional[int] = None):
    """Q[:, 2i] = mean_i - beta*std_i, Q[:, 2i+1] = mean_i + beta*std_i  (gp_opt.py:453-476)."""
    M = inputs.shape[0]
    Q = np.empty((M, 2 * len(gps)), dtype=float)
    step = M if not chunk else int(chunk)
    for i, gp in enumerate(gps):
        for s in range(0, M, max(step, 1)):
            mean, var = gp.predict_noiseless(inputs[s:s + step])
            mean = mean.squeeze(axis=1) if mean.ndim == 2 else mean
            sd = np.sqrt(var.squeeze(axis=1) if var.ndim == 2 else var)
            Q[s:s + step, 2 * i] = mean - beta * sd
            Q[s:s + step, 2 * i + 1] = mean + beta * sd
    return Q


def safe_set(Q, fmin):
    """S = all lower bounds strictly above fmin (gp_opt.py:478-481)."""
    return np.all(Q[:, ::2] > fmin, axis=1)


# --------------------------------------------------------------------------- M, G (a8)
def _is_expander_gp(gps, fmin, beta, x_c, u_c, unsafe_inputs):
    """Refit-based expander test for one candidate (gp_opt.py:579-606)."""
    ok = False
    for i, gp in enumerate(gps):
        if fmin[i] == -np.inf:
            continue
        X0, Y0 = gp.X, gp.Y
        gp.set_XY(np.vstack([X0, x_c]), np.vstack([Y0, u_c[i]]))     # :585-588 -> :227
        mean2, var2 = gp.predict_noiseless(unsafe_inputs)              # :591
        gp.set_XY(gp.X[:-1, :], gp.Y[:-1, :])                          # :594 -> :267
        l2 = mean2.squeeze() - beta * np.sqrt(var2.squeeze())          # :596-598
        ok = bool(np.any(l2 >= fmin[i]))                               # :602
        if not ok:
            break                                                      # :605-606
    return ok


def _is_expander_lipschitz(fmin, lipschitz, x_c, u_c, unsafe_inputs):
    """Lipschitz expander test (gp_opt.py:558-576)."""
    dist = np.sqrt(((unsafe_inputs - x_c[None, :]) ** 2).sum(axis=1))
    ok = False
    for i in range(len(fmin)):
        if fmin[i] == -np.inf:
            continue
        ok = bool(np.any(u_c[i] - lipschitz[i] * dist >= fmin[i]))
        if not ok:
            break
    return ok


def compute_sets(gps, inputs, Q, fmin, beta, scaling, threshold, lipschitz=None,
                 full_sets=False, trace: Optional[dict] = None):
    """Safe set, maximisers M and expanders G (gp_opt.py:483-615).

    Returns (S, M, G).  ``trace`` (optional dict) receives intermediate values used by
    the tests: max_l, max_var, candidate rows in visiting order, number visited."""
    S = safe_set(Q, fmin)                                               # :499
    M = np.zeros_like(S)
    G = np.zeros_like(S)
    if not S.any():                                                     # :504-507
        return S, M, G
    l0, u0 = Q[:, 0], Q[:, 1]
    best_lower = np.max(l0[S])
    M[S] = u0[S] >= best_lower                                          # :511-512
    max_var = np.max(u0[M] - l0[M]) / scaling[0]                        # :513
    lo, up = Q[:, ::2], Q[:, 1::2]                                      # :516-517
    if full_sets:
        s = S.copy()
    else:
        s = S & ~M                                                      # :531
        s[s] = np.max((up[s] - lo[s]) / scaling, axis=1) > max_var      # :534-535
        s[s] = np.any(up[s] - lo[s] > threshold * beta, axis=1)         # :536
    if trace is not None:
        trace.update(max_l=best_lower, max_var=max_var, n_candidates=int(s.sum()))
    if not s.any():                                                     # :538-540
        return S, M, G
    rows = np.flatnonzero(s)
    if full_sets:
        order = np.arange(rows.size)                                    # :555
    else:
        order = np.max(up[s] - lo[s], axis=1).argsort()[::-1]           # :542-552
    unsafe = inputs[~S]
    flags = np.zeros(rows.size, dtype=bool)
    visited = 0
    for idx in order:                                                   # :557
        r = rows[idx]
        visited += 1
        if lipschitz is not None:
            flags[idx] = _is_expander_lipschitz(fmin, lipschitz, inputs[r], up[r], unsafe)
        else:
            flags[idx] = _is_expander_gp(gps, fmin, beta, inputs[r], up[r], unsafe)
        if flags[idx] and not full_sets:                                # :611-612
            break
    G[rows] = flags                                                     # :615
    if trace is not None:
        trace.update(order=rows[order], visited=visited)
    return S, M, G


# --------------------------------------------------------------------------- query (a9, a10)
class NoSafePoints(EnvironmentError):
    pass


def new_query_point(inputs, Q, S, M, G, scaling, ucb=False, num_contexts=0):
    """Masked argmax, first index wins ties (gp_opt.py:617-649)."""
    if not S.any():
        raise NoSafePoints("There are no safe points to evaluate.")
    if ucb:
        row = np.flatnonzero(S)[np.argmax(Q[S, 1])]                     # :634-636
    else:
        MG = M | G                                                      # :642
        width = np.max((Q[MG, 1::2] - Q[MG, ::2]) / scaling, axis=1)    # :643
        row = np.flatnonzero(MG)[np.argmax(width)]                      # :644
    x = inputs[row]
    return (x[:-num_contexts] if num_contexts else x), int(row)


def current_maximum(inputs, Q, S, num_contexts=0):
    """(location, value) of the best safe lower bound, or None (gp_opt.py:677-712)."""
    if not S.any():
        return None
    rows = np.flatnonzero(S)
    k = int(np.argmax(Q[S, 0]))
    x = inputs[rows[k]]
    return (x[:-num_contexts] if num_contexts else x), Q[rows[k], 0]


# --------------------------------------------------------------------------- driver
@dataclass
class GridProblem:
    """State a SafeOpt object carries across ``optimize`` calls (gp_opt.py:347-389)."""
    gps: List
    inputs: np.ndarray
    fmin: np.ndarray
    beta: Callable[[int], float]
    scaling: np.ndarray
    threshold: object = 0
    lipschitz: Optional[np.ndarray] = None
    num_contexts: int = 0
    Q: np.ndarray = field(default=None)
    S: np.ndarray = field(default=None)
    M: np.ndarray = field(default=None)
    G: np.ndarray = field(default=None)

    @classmethod
    def create(cls, gps, parameter_set, fmin, beta=2, threshold=0, lipschitz=None, scaling="auto"):
        gps = list(gps) if isinstance(gps, (list, tuple)) else [gps]
        fm = fmin if isinstance(fmin, list) else [fmin] * len(gps)
        fm = np.atleast_1d(np.asarray(fm, dtype=float).squeeze())                 # :69-72
        bfun = beta if callable(beta) else (lambda t, _b=beta: _b)                  # :74-79
        if isinstance(scaling, str):                                                # :81-84
            zero = np.zeros((1, gps[0].input_dim))
            sc = np.sqrt(np.asarray([g.kern.Kdiag(zero)[0] for g in gps]))
        else:
            sc = np.asarray(scaling, dtype=float)
        lip = None
        if lipschitz is not None:
            lip = lipschitz if isinstance(lipschitz, list) else [lipschitz] * len(gps)
            lip = np.atleast_1d(np.asarray(lip, dtype=float).squeeze())
        return cls(gps=gps, inputs=np.asarray(parameter_set), fmin=fm, beta=bfun, scaling=sc,
                   threshold=threshold, lipschitz=lip)

    @property
    def t(self):
        return self.gps[0].X.shape[0]

    def optimize(self, ucb=False, chunk=None, trace=None):
        """One ``SafeOpt.optimize`` (gp_opt.py:651-675); returns (x_next, row)."""
        b = self.beta(self.t)
        self.Q = confidence_intervals(self.gps, self.inputs, b, chunk=chunk)
        if ucb:
            self.S = safe_set(self.Q, self.fmin)
            self.M = np.zeros_like(self.S)
            self.G = np.zeros_like(self.S)
        else:
            self.S, self.M, self.G = compute_sets(self.gps, self.inputs, self.Q, self.fmin, b,
                                                  self.scaling, self.threshold, self.lipschitz, trace=trace)
        return new_query_point(self.inputs, self.Q, self.S, self.M, self.G, self.scaling, ucb=ucb)

    def add_new_data_point(self, x, y):
        """gp_opt.py:230-255 (NaN entries skip that GP)."""
        x = np.atleast_2d(x)
        y = np.atleast_2d(y)
        for i, gp in enumerate(self.gps):
            keep = ~np.isnan(y[:, i])
            if keep.any():
                gp.set_XY(np.vstack([gp.X, x[keep]]), np.vstack([gp.Y, y[keep][:, [i]]]))


# --------------------------------------------------------------------------- swarm (a12-a14)
def penalty(slack):
    """Piecewise constraint-violation penalty (gp_opt.py:874-899)."""
    slack = np.atleast_1d(np.asarray(slack, dtype=float))
    p = np.minimum(slack, 0.0)
    p = np.where((slack < 0) & (slack > -0.001), 2 * p, p)
    p = np.where((slack <= -0.001) & (slack > -0.1), 5 * p, p)
    p = np.where((slack <= -0.1) & (slack > -1), 10 * p, p)
    big = slack < -1
    p = np.where(big, -300.0 * np.minimum(slack, 0.0) ** 2, p)
    return p


def particle_fitness(gps, fmin, beta, scaling, swarm_type, particles, best_lower_bound=-np.inf):
    """Fitness value and safety flag of every particle (gp_opt.py:901-1013)."""
    mean, var = gps[0].predict_noiseless(particles)
    mean = mean.squeeze(axis=1)
    sd = np.sqrt(var.squeeze(axis=1))
    lower = mean - beta * sd
    upper = mean + beta * sd
    P = particles.shape[0]
    if swarm_type == "greedy":                                           # :938-939
        return lower, np.ones(P, dtype=bool)
    values = sd / scaling[0]                                             # :943
    if swarm_type == "safe_set":
        interest = None
    elif swarm_type == "expanders":
        interest = len(gps) * np.ones(P)                                 # :956-957
    elif swarm_type == "maximizers":
        interest = expit(10 * (upper - best_lower_bound) / scaling[0])   # :959-960
    else:
        raise AssertionError("Invalid swarm type")
    safe = np.ones(P, dtype=bool)
    total_pen = np.zeros(P)
    for i, gp in enumerate(gps):                                         # :969
        if i > 0:
            mean, var = gp.predict_noiseless(particles)
            sd = np.sqrt(var.squeeze(axis=1))
            lower = mean.squeeze(axis=1) - beta * sd
            values = np.maximum(values, sd / scaling[i])                 # :978
        if fmin[i] == -np.inf:
            continue
        slack = lower - fmin[i]
        safe &= slack >= 0                                               # :987
        if swarm_type == "safe_set":
            continue
        slack = slack / scaling[i]                                       # :994
        total_pen += penalty(slack)
        if swarm_type == "expanders":
            interest = interest * norm.pdf(slack, scale=0.2)             # :1000
    if swarm_type == "safe_set":
        return lower, safe                                               # :1004-1005
    return (values + total_pen) * interest, safe                         # :1008-1013


def pso_step(positions, velocities, best_positions, best_values, global_best, r1, r2,
             inertia, velocity_scale, bounds, fitness):
    """One particle-swarm iteration with host-supplied randoms (safeopt/swarm.py:98-146)."""
    velocities = velocities * inertia + (r1 * (best_positions - positions)
                                         + r2 * (global_best - positions)) / velocity_scale
    vmax = 10 * velocity_scale
    velocities = np.clip(velocities, -vmax, vmax)
    positions = positions + velocities
    if bounds is not None:
        positions = np.clip(positions, bounds[:, 0], bounds[:, 1])
    values, safe = fitness(positions)
    upd = (values > best_values) & safe
    best_values = np.where(upd, values, best_values)
    best_positions = np.where(upd[:, None], positions, best_positions)
    global_best = best_positions[np.argmax(best_values)].copy()
    return positions, velocities, best_positions, best_values, global_best


def select_new_safe_points(kern, S, best_positions, scaling0, limit=0.95):
    """Swarm safe-set insertion (gp_opt.py:1088-1110): particle j joins iff its prior correlation
    ``k(x_j, .) / scaling0**2`` with every old safe point and every particle accepted before it is
    ``<= limit``.  Returns (accepted mask over the particles, min |corr - limit| over the pairs compared).

    Row-at-a-time restatement (the reference builds the full P x (|S|+P) matrix first); the comparison
    set and order are the reference's."""
    S = np.asarray(S, dtype=float)
    best_positions = np.asarray(best_positions, dtype=float)
    accepted = np.zeros(best_positions.shape[0], dtype=bool)
    margin = np.inf
    for j in range(best_positions.shape[0]):
        pool = np.vstack((S, best_positions[accepted]))
        corr = kern.K(best_positions[[j]], pool)[0] / scaling0 ** 2
        margin = min(margin, float(np.min(np.abs(corr - limit)))) if corr.size else margin
        if np.all(corr <= limit):
            accepted[j] = True
    return accepted, margin
