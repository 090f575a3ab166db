"""SafeOpt / SafeOptSwarm with the reference's class surface, executed on B200 GPUs.

Drop-in for the hot path of /root/reference/safeopt/gp_opt.py: same constructor signatures,
methods, attributes and exceptions; the arithmetic below ``update_confidence_intervals`` /
``compute_sets`` / ``get_new_query_point`` / ``_compute_particle_fitness`` runs in the
hand-written sm_100a kernels behind include/safeopt_b200.h.  Host Python only orchestrates:
it reads hyper-parameters off the GPy-style models, launches stream-ordered kernels, exchanges
64-byte records between ranks and maps row indices back to parameters.

Multi-GPU: when ``torch.distributed`` is initialised every rank constructs the same optimiser;
rows are sharded in contiguous blocks (distributed.shard_bounds), the fit is recomputed per rank.
"""
from __future__ import annotations

import logging
import os
from collections.abc import Sequence
from functools import partial
from typing import List, Optional

import numpy as np

from . import _lib
from .distributed import (Comm, combine_max_first, gather_padded_rows, gather_ragged, gather_row_blocks, order_candidates,
                          reduce_max_records, reduce_safe_records, shard_bounds)
from .engine import MAX_REC_DTYPE, SAFE_REC_DTYPE, DeviceEngine
from .gpmodel import extract_hyper, fingerprint
from .swarm import DeviceSwarm, SwarmOptimization
from .utilities import detect_grid, grid_rows_from_index, linearly_spaced_combinations

__all__ = ["SafeOpt", "SafeOptSwarm", "GaussianProcessOptimization"]

import struct as _struct
_SAFE_STRUCT = _struct.Struct("<qdqdq")       # so_safe_record: n_safe, max_l0, argmax_l0, max_u0, argmax_u0
_MAX_STRUCT = _struct.Struct("<qddq")         # so_max_record: n_max, max_width0, best_value, best_row


class GaussianProcessOptimization(object):
    """Common bookkeeping of the optimisers (reference: gp_opt.py:30-278).

    Parameters
    ----------
    gp : GPy-style Gaussian process, or a list of them (first = objective, rest = constraints).
    fmin : float or list of floats -- safety thresholds (``-np.inf`` / ``None`` = unconstrained).
    beta : float or callable(t) -- confidence-interval scaling.
    num_contexts : int -- trailing input columns that are contexts.
    threshold : float or list -- expanders narrower than this (unscaled) are ignored.
    scaling : list of floats or "auto" (sqrt of each kernel's prior variance).
    """

    def __init__(self, gp, fmin, beta=2, num_contexts=0, threshold=0, scaling="auto"):
        super(GaussianProcessOptimization, self).__init__()
        self.gps = gp if isinstance(gp, list) else [gp]
        self.gp = self.gps[0]

        self.fmin = fmin
        if not isinstance(self.fmin, list):
            self.fmin = [self.fmin] * len(self.gps)
        # None = unconstrained, like -inf (np.asarray would turn None into NaN, and every comparison with NaN is False)
        self.fmin = [(-np.inf if f is None else f) for f in self.fmin]
        self.fmin = np.atleast_1d(np.asarray(self.fmin, dtype=float).squeeze())
        if np.isnan(self.fmin).any():
            raise ValueError("fmin must not contain NaN (use -np.inf or None for an unconstrained GP)")

        self.beta = beta if callable(beta) else (lambda t: beta)

        if isinstance(scaling, str) and scaling == "auto":
            origin = np.zeros((1, self.gps[0].input_dim))
            self.scaling = np.sqrt(np.asarray([g.kern.Kdiag(origin)[0] for g in self.gps], dtype=float))
        else:
            self.scaling = np.asarray(scaling, dtype=float)
            if self.scaling.shape[0] != len(self.gps):
                raise ValueError("The number of scaling values should be equal to the number of GPs")

        self.threshold = threshold
        self._parameter_set = None
        self.bounds = None
        self.num_samples = 0
        self.num_contexts = num_contexts

        self._x = None
        self._y = None
        self._get_initial_xy()

    # ---- data access (gp_opt.py:101-130)
    @property
    def x(self):
        return self._x

    @property
    def y(self):
        return self._y

    @property
    def data(self):
        """The measurements held by the GP models, ``(x, y)``."""
        return self._x, self._y

    @property
    def t(self):
        """Time step = number of measurements."""
        return self._x.shape[0]

    def _get_initial_xy(self):
        self._x = self.gp.X
        cols = [self.gp.Y]
        for other in self.gps[1:]:
            if not np.allclose(self._x, other.X):
                raise NotImplementedError("The GPs have different measurements.")
            cols.append(other.Y)
        self._y = np.concatenate(cols, axis=1)

    def plot(self, *args, **kwargs):
        """Plotting is outside the accelerated path (matplotlib-only code in the reference)."""
        raise NotImplementedError("plotting is out of scope for safeopt_b200; use the reference's plot helpers "
                                  "with opt.Q / opt.S / opt.M / opt.G")

    # ---- data management (gp_opt.py:187-278)
    def _add_context(self, x, context):
        context = np.atleast_2d(context)
        nc = context.shape[1]
        out = np.empty((x.shape[0], x.shape[1] + nc), dtype=float)
        out[:, :x.shape[1]] = x
        out[:, x.shape[1]:] = context
        return out

    def _add_data_point(self, gp, x, y, context=None):
        """Append one observation to a single GP (does not touch ``self.x`` / ``self.y``)."""
        if context is not None:
            x = self._add_context(x, context)
        gp.set_XY(np.vstack([gp.X, x]), np.vstack([gp.Y, y]))

    def add_new_data_point(self, x, y, context=None):
        """Add an observation ``y`` (one column per GP, NaN = not measured) at parameters ``x``."""
        x = np.atleast_2d(x)
        y = np.atleast_2d(y)
        if self.num_contexts:
            x = self._add_context(x, context)
        for i, gp in enumerate(self.gps):
            ok = ~np.isnan(y[:, i])
            if np.any(ok):
                self._add_data_point(gp, x[ok, :], y[ok, [i]])
        self._x = np.concatenate((self._x, x), axis=0)
        self._y = np.concatenate((self._y, y), axis=0)

    def _remove_last_data_point(self, gp):
        gp.set_XY(gp.X[:-1, :], gp.Y[:-1, :])

    def remove_last_data_point(self):
        """Undo the last ``add_new_data_point``."""
        last_y = self._y[-1]
        for gp, yi in zip(self.gps, last_y):
            if not np.isnan(yi):
                gp.set_XY(gp.X[:-1, :], gp.Y[:-1, :])
        self._x = self._x[:-1, :]
        self._y = self._y[:-1, :]


def _raw_bytes(a):
    return a.tobytes() if hasattr(a, "tobytes") else repr(a)


def _quick_signature(gp):
    """(X object, Y object, kernel object, raw parameter bytes) of a plain stationary-kernel model, or None when the model
    does not have that shape (composite kernels take the full extract_hyper path every time)."""
    try:
        k = gp.kern
        return (gp.X, gp.Y, k, (_raw_bytes(k.variance), _raw_bytes(k.lengthscale), _raw_bytes(gp.likelihood.variance),
                                getattr(gp, "mean_function", None) is None, getattr(gp, "normalizer", None) in (None, False)))
    except AttributeError:
        return None


class _DeviceFits:
    """Keeps the device-side fit of each GP in step with the user's model objects.

    The user's objects stay the source of truth (``set_XY`` is still called on them); before
    every device pass their data and hyper-parameters are fingerprinted.  A change that is exactly "one row
    appended" or "last row removed" under unchanged hyper-parameters -- what ``add_new_data_point`` /
    ``remove_last_data_point`` do (gp_opt.py:230-278) -- is applied to the resident factorisation in O(N^2)
    (``so_fit_append`` / ``so_fit_remove_last``); anything else refits from scratch (``so_fit``).
    ``SAFEOPT_B200_INCREMENTAL_FIT=0`` disables the incremental path."""

    def __init__(self, engine: DeviceEngine, gps):
        self.engine = engine
        self.gps = gps
        self._fp = [None] * len(gps)
        self._ident = [None] * len(gps)       # (X object, Y object, scalar hyper-parameters) seen at the last refresh
        self._data = [None] * len(gps)        # (X, Y, hyper key) the device fit was built from
        self.hypers = [None] * len(gps)
        self.async_fit = False                # set by an owner that checks engine.check_fits() after its next synchronisation
        self.generation = 0                   # bumped whenever a device fit changed
        self.refits = 0
        self.copies = 0                       # refits served by copying a twin's factorisation (so_fit_like)
        self.appends = 0
        self.removals = 0
        self.incremental = os.environ.get("SAFEOPT_B200_INCREMENTAL_FIT", "1") != "0"
        # GPs with identical inputs, kernel and noise share K(X,X), its factor and every kernel row: they are evaluated by
        # one launch (one contraction, one V.z per GP).  SAFEOPT_B200_SHARE_FITS=0 evaluates every GP on its own.
        self.share = os.environ.get("SAFEOPT_B200_SHARE_FITS", "1") != "0"
        self.groups = [[i] for i in range(len(gps))]

    def refresh(self, after_fit=None):
        changed = False
        for i, gp in enumerate(self.gps):
            # Fast path (this runs before every K2 launch; on small grids the step is host-bound): the same X / Y / kernel
            # OBJECTS with byte-identical parameter values mean nothing changed -- GPy's own posterior cache is keyed on
            # set_XY / parameter updates too, not on array contents.  Parameter VALUES are compared (GPy's optimiser
            # updates them in place), data by identity (set_XY replaces the arrays).
            sig = _quick_signature(gp)
            last = self._ident[i]
            if sig is not None and last is not None and last[0] is sig[0] and last[1] is sig[1] and last[2] is sig[2] \
                    and last[3] == sig[3] and self.hypers[i] is not None:
                continue
            hyper = extract_hyper(gp)
            ident = sig
            fp = fingerprint(gp, hyper)
            if fp == self._fp[i]:
                self._ident[i] = ident
                continue
            X = np.ascontiguousarray(np.asarray(gp.X, dtype=float))
            Y = np.ascontiguousarray(np.asarray(gp.Y, dtype=float)[:, 0])
            key = (hyper.kind, hyper.lengthscale.tobytes(), hyper.variance, hyper.noise_var)
            done = False
            prev = self._data[i]
            if self.incremental and prev is not None and prev[2] == key and prev[0].shape[1] == X.shape[1]:
                Xp, Yp = prev[0], prev[1]
                if X.shape[0] == Xp.shape[0] + 1 and np.array_equal(X[:-1], Xp) and np.array_equal(Y[:-1], Yp):
                    done = self.engine.fit_append(i, X[-1], Y[-1])
                    self.appends += int(done)
                elif X.shape[0] == Xp.shape[0] - 1 and X.shape[0] >= 1 and np.array_equal(X, Xp[:-1]) \
                        and np.array_equal(Y, Yp[:-1]):
                    self.engine.fit_remove_last(i)
                    self.removals += 1
                    done = True
            if not done:
                # a GP fitted earlier in this pass (or still current) on the same inputs, kernel and noise shares K, L, L^-1
                twin = None
                if self.share and hasattr(self.engine, "fit_like"):
                    for j in range(len(self.gps)):
                        dj = self._data[j]
                        if j != i and dj is not None and self._fp[j] is not None and dj[2] == key and dj[0].shape == X.shape \
                                and np.array_equal(dj[0], X):
                            twin = j
                            break
                if twin is not None:
                    self.engine.fit_like(i, twin, Y)
                    self.copies += 1
                elif self.async_fit and hasattr(self.engine, "check_fits"):
                    # no host wait: table build and posterior are queued behind the factorisation; a failed Cholesky is reported
                    # when the records of the set pass arrive (the owner calls engine.check_fits() there)
                    self.engine.fit(i, X, Y, hyper.kind, hyper.lengthscale, hyper.variance, hyper.noise_var, wait=False)
                else:
                    self.engine.fit(i, X, Y, hyper.kind, hyper.lengthscale, hyper.variance, hyper.noise_var)
                self.refits += 1
            changed = True
            self._fp[i] = fp
            self._ident[i] = ident
            self._data[i] = (X.copy(), Y.copy(), key)
            self.hypers[i] = hyper
            if after_fit is not None:
                after_fit(i, hyper)
        if changed:
            self.generation += 1
            self._regroup()

    MAX_GROUP = 4       # kMaxOut of the kernels

    def _regroup(self):
        """Partition the GPs into groups that share (X, kernel, noise); order of first appearance, at most MAX_GROUP each."""
        groups, index = [], {}
        for i, data in enumerate(self._data):
            key = None if (data is None or not self.share) else (data[0].shape, data[0].tobytes(), data[2])
            slot = index.get(key) if key is not None else None
            if slot is None or len(groups[slot]) >= self.MAX_GROUP:
                if key is not None:
                    index[key] = len(groups)
                groups.append([i])
            else:
                groups[slot].append(i)
        self.groups = groups

    def invalidate(self):
        """Forget the resident fits: the next refresh refits every GP from scratch."""
        self._fp = [None] * len(self.gps)
        self._ident = [None] * len(self.gps)
        self._data = [None] * len(self.gps)


class SafeOpt(GaussianProcessOptimization):
    """Safe Bayesian optimisation over a finite parameter set (reference: gp_opt.py:281-712).

    Parameters are the reference's (``gp, parameter_set, fmin, lipschitz=None, beta=2,
    num_contexts=0, threshold=0, scaling='auto'``) plus ``device`` (CUDA device, default current),
    ``distributed`` (default True: shard the rows over the ranks of ``torch.distributed`` when it is
    initialised; False: this object evaluates every row on its own GPU) and ``precision``: ``'fp64'``
    (default, the reference's arithmetic: bounds within 1e-9, masks bit-exact) or ``'fp32'`` -- the
    posterior contraction on the tcgen05 tensor cores in error-compensated TF32 (posterior within 1e-4
    relative, masks exact outside that band around the thresholds; product grids with RBF kernels and
    N <= 256, anything else keeps the fp64 kernels and says so once); fit and set logic stay fp64.

    Examples
    --------
    >>> from safeopt_b200 import SafeOpt, linearly_spaced_combinations, GPRegression
    >>> import numpy as np
    >>> gp = GPRegression(np.array([[0.]]), np.array([[1.]]), noise_var=0.01 ** 2)
    >>> parameter_set = linearly_spaced_combinations([[-1., 1.]], num_samples=100)
    >>> opt = SafeOpt(gp, parameter_set, fmin=[0.])
    >>> next_parameters = opt.optimize()                                   # doctest: +SKIP
    >>> opt.add_new_data_point(next_parameters, np.array([[1.]]))          # doctest: +SKIP
    """

    F32_MAX_N = 256          # the fp32 tensor-core kernel keeps N <= 256 (TMEM accumulator columns)
    precision = "fp64"       # arithmetic of the posterior contraction (constructor keyword)

    def __init__(self, gp, parameter_set, fmin, lipschitz=None, beta=2, num_contexts=0, threshold=0,
                 scaling="auto", device=None, distributed=True, precision="fp64"):
        super(SafeOpt, self).__init__(gp, fmin=fmin, beta=beta, num_contexts=num_contexts, threshold=threshold,
                                      scaling=scaling)
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        self.precision = precision
        self._f32_warned = False
        parameter_set = np.asarray(parameter_set, dtype=float)
        # SAFEOPT_B200_GRID_FAST_PATH=0 forces the explicit-rows kernels (tests / A-B measurements)
        axes = None
        if os.environ.get("SAFEOPT_B200_GRID_FAST_PATH", "1") != "0" and parameter_set.shape[1] + self.num_contexts <= 6:
            axes = detect_grid(parameter_set)
        self._param_axes = axes
        self._context_on_device = None
        if self.num_contexts > 0:
            zeros = np.zeros((parameter_set.shape[0], self.num_contexts), dtype=parameter_set.dtype)
            self.inputs = np.hstack((parameter_set, zeros))
            self.parameter_set = self.inputs[:, :-self.num_contexts]
            if axes is not None:
                # contexts are constant over the candidates (gp_opt.py:424-451): to the device they are grid axes with ONE
                # point each, appended after the parameter axes (the row order does not change), so a context problem keeps
                # the separable-table kernels -- k_ctx(c, c_n) ends up folded into the per-training-point table entries
                axes = list(axes) + [np.zeros(1) for _ in range(self.num_contexts)]
        elif axes is not None:
            # a verified product grid: bounds and num_samples (gp_opt.py:414-422) follow from the axes, without
            # the per-column np.unique sorts over all M rows
            self.inputs = self._parameter_set = parameter_set
            self.bounds = [(np.min(a), np.max(a)) for a in axes]
            self.num_samples = [len(a) for a in axes]
        else:
            self.inputs = self.parameter_set = parameter_set

        self.liptschitz = lipschitz          # (sic) attribute name of the reference, gp_opt.py:366
        if self.liptschitz is not None:
            if not isinstance(self.liptschitz, list):
                self.liptschitz = [self.liptschitz] * len(self.gps)
            self.liptschitz = np.atleast_1d(np.asarray(self.liptschitz, dtype=float).squeeze())
        self._use_lipschitz = lipschitz is not None

        # ---- device state
        self._engine = DeviceEngine(device, max_gps=len(self.gps))
        # NCCL staging tensors live on the engine's device, not the current one; distributed=False keeps this object
        # rank-local although torch.distributed is initialised (it then evaluates ALL rows on this GPU)
        self._comm = Comm(self._engine.device, enabled=distributed)
        self._fits = _DeviceFits(self._engine, self.gps)
        self._fits.async_fit = os.environ.get("SAFEOPT_B200_ASYNC_FIT", "1") != "0"
        # cross-rank records through peer-mapped memory written by the kernels themselves (so_xchg_*); where IPC mapping is
        # not available -- and in the CPU tests' stand-in engine -- records go through all-gathers between chained kernels
        self._peer = bool(hasattr(self._engine, "connect_exchange") and self._engine.connect_exchange(self._comm))
        self._fused = (hasattr(self._engine, "sets_fused") and (self._peer or not self._comm.active)
                       and os.environ.get("SAFEOPT_B200_FUSED_SETS", "1") != "0")
        n_rows = self.inputs.shape[0]
        self._row0, self._row1 = shard_bounds(n_rows, self._comm.world, self._comm.rank)
        m_local = self._row1 - self._row0
        self._grid_axes = axes
        self._rows_d = None
        self._grid_state = {}                 # GP index -> which grid tables are current (see _ensure_grid_tables)
        if self._grid_axes is not None:
            self._engine.define_grid(self._grid_axes)
        else:
            self._rows_d = self._engine.to_device(np.ascontiguousarray(self.inputs[self._row0:self._row1]))
        eng = self._engine
        G = len(self.gps)
        self._Q_d = eng.empty((m_local, 2 * G))
        self._S_d = eng.zeros((m_local,), "u8")
        self._M_d = eng.zeros((m_local,), "u8")
        self._mean_d = eng.empty((G, m_local))
        self._var_d = eng.empty((G, m_local))
        self._rec_safe_d = eng.zeros((64,), "u8")
        self._rec_max_d = eng.zeros((64,), "u8")
        self._n_cand_d = eng.zeros((1,), "i64")
        self._cand_key_d = None
        self._cand_row_d = None
        self._grid_strides = None
        self._thr_cache = None
        self._k2_tape = None                  # (key, engine tape) of the last update_confidence_intervals
        self._tables_version = 0
        self._G_rows: List[int] = []          # global rows currently in the expander set
        self._host_cache = {}
        self._safe_info = None                # combined record of the last compute_safe_set
        self._max_info = None
        self._ci_beta = None
        self.last_trace = {}

    # ------------------------------------------------------------------ reference properties
    @property
    def use_lipschitz(self):
        """Whether expanders are judged with the Lipschitz constant instead of the GP (gp_opt.py:391-407)."""
        return self._use_lipschitz

    @use_lipschitz.setter
    def use_lipschitz(self, value):
        if value and self.liptschitz is None:
            raise ValueError("Lipschitz constant not defined")
        self._use_lipschitz = value

    @property
    def parameter_set(self):
        """Discrete parameter samples for Bayesian optimisation."""
        return self._parameter_set

    @parameter_set.setter
    def parameter_set(self, parameter_set):
        self._parameter_set = parameter_set
        self.bounds = list(zip(np.min(self._parameter_set, axis=0), np.max(self._parameter_set, axis=0)))
        self.num_samples = [len(np.unique(self._parameter_set[:, i])) for i in range(self._parameter_set.shape[1])]

    @property
    def context_fixed_inputs(self):
        """Fixed inputs for the current context (gp_opt.py:424-431)."""
        n = self.gp.input_dim - 1
        nc = self.num_contexts
        if nc > 0:
            contexts = self.inputs[0, -self.num_contexts:]
            return list(zip(range(n, n - nc, -1), contexts))

    @property
    def context(self):
        """Current context variables."""
        if self.num_contexts:
            return self.inputs[0, -self.num_contexts:]

    @context.setter
    def context(self, context):
        if self.num_contexts:
            if context is None:
                raise ValueError("Need to provide value for context.")
            self.inputs[:, -self.num_contexts:] = context
            ctx = np.atleast_1d(np.asarray(context, dtype=float)).ravel()
            if self._rows_d is not None:
                self._rows_d[:, -self.num_contexts:] = self._engine.to_device(ctx)     # same buffer: a taped launch stays valid
            if self._grid_axes is not None and (self._context_on_device is None or not np.array_equal(ctx, self._context_on_device)):
                # a new context = new one-point axes: the grid description and every GP's tables are rebuilt
                self._grid_axes = list(self._param_axes) + [np.array([c], dtype=float) for c in ctx]
                self._engine.define_grid(self._grid_axes)
                self._grid_state.clear()
                self._tables_version += 1
                self._context_on_device = ctx.copy()

    # ------------------------------------------------------------------ host views of device state
    def _gather_rows(self, local: np.ndarray) -> np.ndarray:
        return gather_row_blocks(self._comm, local, self.inputs.shape[0])

    def _check_fits(self):
        """After a host wait: did a factorisation started without waiting (so_fit_async) fail?"""
        if getattr(self._engine, "_pending_fits", None):
            self._engine.check_fits()

    def _host(self, key, tensor, as_bool=False):
        if key not in self._host_cache:
            arr = tensor.cpu().numpy()
            self._check_fits()
            arr = self._gather_rows(arr)
            self._host_cache[key] = arr.astype(bool) if as_bool else arr
        return self._host_cache[key]

    def local_block(self, name):
        """Rank-local shard of ``'Q'``, ``'S'`` or ``'M'`` as a host array together with its global row range
        ``(row0, row1)``.  Never communicates -- the accessor to use from ONE rank (logging, plotting) in a multi-GPU run,
        where the properties ``Q`` / ``S`` / ``M`` are collectives."""
        t = {"Q": self._Q_d, "S": self._S_d, "M": self._M_d}[name]
        arr = t.cpu().numpy()
        return (arr if name == "Q" else arr.astype(bool)), (self._row0, self._row1)

    @property
    def Q(self):
        """Confidence intervals, ``(M, 2G)`` float64: columns ``l0, u0, l1, u1, ...`` (copied from the device on access).

        With ``torch.distributed`` initialised, ``Q``, ``S`` and ``M`` reassemble the row-sharded arrays with an
        all-gather: they are COLLECTIVES that every rank must read (the result is cached until the next device pass);
        reading one on a single rank only deadlocks the job -- use :meth:`local_block` there."""
        return self._host("Q", self._Q_d)

    @property
    def S(self):
        """Safe set mask ``(M,)`` bool (multi-rank: a collective, see :attr:`Q`)."""
        return self._host("S", self._S_d, as_bool=True)

    @property
    def M(self):
        """Maximiser mask ``(M,)`` bool (multi-rank: a collective, see :attr:`Q`)."""
        return self._host("M", self._M_d, as_bool=True)

    @property
    def G(self):
        """Expander mask ``(M,)`` bool."""
        g = np.zeros(self.inputs.shape[0], dtype=bool)
        if self._G_rows:
            g[np.asarray(self._G_rows, dtype=np.int64)] = True
        return g

    def _invalidate_host(self, *keys):
        for k in (keys or list(self._host_cache)):
            self._host_cache.pop(k, None)

    def _row_point(self, row: int) -> np.ndarray:
        """Parameters (and contexts) of one global row -- scalar arithmetic, this runs once per optimize()."""
        if self._grid_axes is not None:
            if self._grid_strides is None:
                from .utilities import grid_row_strides
                self._grid_strides = grid_row_strides([len(a) for a in self._grid_axes])
            return np.array([a[(row // st) % len(a)] for a, st in zip(self._grid_axes, self._grid_strides)], dtype=float)
        return np.asarray(self.inputs[row], dtype=float)

    def _row_coordinates(self, rows) -> np.ndarray:
        rows = np.atleast_1d(np.asarray(rows, dtype=np.int64))
        if self._grid_axes is not None:
            return grid_rows_from_index(self._grid_axes, rows)
        return np.asarray(self.inputs[rows], dtype=float)

    # ------------------------------------------------------------------ hot path
    def _after_fit(self, i, hyper):
        # the grid tables of GP i are rebuilt lazily, by whoever needs them next (_ensure_grid_tables): only the GP that
        # provides a group's factorisation needs the large operand tables, the others only the small per-axis tables and
        # only if an expander search runs
        self._grid_state[i] = None

    def _use_grid_kernel(self, i) -> bool:
        return self._grid_axes is not None and self._fits.hypers[i].kind == _lib.KERNEL_RBF

    def _use_f32(self, i) -> bool:
        if self.precision != "fp32":
            return False
        ok = self._use_grid_kernel(i) and self.gps[i].X.shape[0] <= self.F32_MAX_N and hasattr(self._engine, "posterior_grid_f32")
        if not ok and not self._f32_warned:
            self._f32_warned = True
            logging.warning("precision='fp32' needs a product grid, an RBF kernel and N <= %d: using the fp64 kernels" % self.F32_MAX_N)
        return ok

    def _ensure_grid_tables(self, i, level):
        """Grid tables of GP ``i`` after its last (re)fit.  level 'axes': per-axis tables only (expander kernel);
        'fp64': + the scaled-operand table of this rank's rows (fp64 grid kernel); 'fp32': + the TF32 operand planes."""
        have = self._grid_state.get(i) or set()
        eng, m_local = self._engine, self._row1 - self._row0
        if level == "fp64" and "fp64" not in have:
            eng.prepare_grid(i, self._row0, m_local)
            have |= {"axes", "fp64"}
        elif "axes" not in have:
            eng.prepare_grid(i, self._row0, 0)          # per-axis and product tables, no scaled-operand table
            have.add("axes")
        if level == "fp32" and "fp32" not in have:
            eng.prepare_grid_f32(i, self._row0, m_local)
            have.add("fp32")
        self._grid_state[i] = have

    def _ensure_rows_on_device(self):
        """Explicit rows are needed when some GP cannot use the separable grid tables."""
        if self._rows_d is None:
            self._rows_d = self._engine.grid_rows(self._row0, self._row1 - self._row0)
        return self._rows_d

    def update_confidence_intervals(self, context=None):
        """Recompute ``Q`` from the GP posteriors on every candidate (reference: gp_opt.py:453-476).

        One fused kernel per GP: kernel rows, the L^-1 contraction on the fp64 tensor pipe,
        mean/variance, ``l/u`` and that GP's safe bit (the AND over GPs of gp_opt.py:481)."""
        beta = self.beta(self.t)
        self.context = context
        self._fits.refresh(self._after_fit)
        eng = self._engine
        # unchanged fits, tables, thresholds and beta: re-issue the remembered launches (engine tape) -- nothing to rebuild
        key = (self._fits.generation, self._tables_version, beta, self.fmin.tobytes(), self.precision)
        if self._k2_tape is not None and self._k2_tape[0] == key:
            eng.replay_tape(self._k2_tape[1])
            self._ci_beta = beta
            self._safe_info = None
            if self._host_cache:
                self._invalidate_host()
            return
        taping = hasattr(eng, "start_tape")
        if taping:
            eng.start_tape()
        m_local = self._row1 - self._row0
        first = True
        for group in self._fits.groups:
            # GPs that share data, kernel and noise go through one launch; the S bit is the AND over all GPs (gp_opt.py:481)
            rows_arg = None if self._use_grid_kernel(group[0]) else self._ensure_rows_on_device()
            if rows_arg is None and self._use_f32(group[0]):
                # fp32 arithmetic mode: tcgen05 tensor cores (3xTF32), TMEM accumulators; fp64 epilogue
                self._ensure_grid_tables(group[0], "fp32")
                eng.posterior_grid_f32(group, self._row0, m_local, beta, [self.fmin[i] for i in group],
                                       means=[self._mean_d[i] for i in group], variances=[self._var_d[i] for i in group],
                                       Q=self._Q_d, q_cols=[2 * i for i in group], S=self._S_d,
                                       safe_mode=_lib.SAFE_WRITE if first else _lib.SAFE_AND)
                first = False
                continue
            if rows_arg is None:
                self._ensure_grid_tables(group[0], "fp64")
            done = False
            if len(group) > 1:
                done = eng.posterior_multi(group, rows_arg, self._row0, m_local, beta, [self.fmin[i] for i in group],
                                           means=[self._mean_d[i] for i in group], variances=[self._var_d[i] for i in group],
                                           Q=self._Q_d, q_cols=[2 * i for i in group], S=self._S_d,
                                           safe_mode=_lib.SAFE_WRITE if first else _lib.SAFE_AND)
                first = first and not done
            if not done:
                for i in group:
                    mode = _lib.SAFE_WRITE if first else _lib.SAFE_AND
                    first = False
                    if rows_arg is None:
                        self._ensure_grid_tables(i, "fp64")
                        eng.posterior_grid(i, self._row0, m_local, beta, self.fmin[i], mean=self._mean_d[i], var=self._var_d[i],
                                           Q=self._Q_d, q_col=2 * i, S=self._S_d, safe_mode=mode)
                    else:
                        eng.posterior_rows(i, rows_arg, beta, self.fmin[i], mean=self._mean_d[i], var=self._var_d[i],
                                           Q=self._Q_d, q_col=2 * i, S=self._S_d, safe_mode=mode)
        self._k2_tape = (key, eng.stop_tape()) if taping else None
        self._ci_beta = beta
        self._ci_fmin = self.fmin.copy()
        self._safe_info = None
        self._invalidate_host()

    def compute_safe_set(self):
        """Safe set from the current bounds (reference: gp_opt.py:478-481).

        The S bytes were written by ``update_confidence_intervals``; this pass reduces them to the
        record the later steps need (count, best safe lower/upper bound and where)."""
        if self._ci_beta is None:
            raise RuntimeError("call update_confidence_intervals() first")
        if not np.array_equal(self._ci_fmin, self.fmin):
            # thresholds changed since the bounds were computed: redo the (fused) pass
            self.update_confidence_intervals(context=self.context)
        eng = self._engine
        if self._fused:
            # the fused set kernel without its candidate pass: the safe record of every rank arrives through the peer buffers
            # (no collective); it also rewrites the maximiser mask, which is a function of Q and S only
            G, world = len(self.gps), self._comm.world
            host = eng.sets_fused(self._Q_d, G, self._row0, self._S_d, self.scaling, np.zeros(G), False, self._M_d, None, None)
            self._safe_info = reduce_safe_records(host[:world * 64].view(SAFE_REC_DTYPE).reshape(-1))
            self._invalidate_host("M")
        else:
            eng.reduce_safe(self._Q_d, len(self.gps), self._row0, self._S_d, self._rec_safe_d)
            self._safe_info = reduce_safe_records(self._comm.gather_records(self._rec_safe_d, SAFE_REC_DTYPE))
        self._check_fits()
        self._invalidate_host("S")

    def _record_buffers(self):
        """Device buffer holding every rank's records of the three set passes back to back
        ([world x safe record][world x max record][world x candidate count]) so the host needs one copy.
        On a single GPU the kernels write straight into it; with several ranks each pass is followed by a
        stream-ordered all-gather of the 64-byte record."""
        if getattr(self, "_recs_all_d", None) is None:
            eng, world, t = self._engine, self._comm.world, self._engine.torch
            self._recs_all_d = eng.zeros((world * 136,), "u8")
            self._safe_all_d = self._recs_all_d[:world * 64].view(world, 64)
            self._max_all_d = self._recs_all_d[world * 64:world * 128].view(world, 64)
            self._ncand_all_d = self._recs_all_d[world * 128:].view(t.int64)
            if world == 1:
                self._rec_safe_l, self._rec_max_l, self._n_cand_l = self._safe_all_d[0], self._max_all_d[0], self._ncand_all_d
            else:
                self._rec_safe_l, self._rec_max_l, self._n_cand_l = self._rec_safe_d, self._rec_max_d, self._n_cand_d
        return self._recs_all_d

    def _share(self, all_d, local_d):
        if self._comm.active:
            self._comm.all_gather_into(all_d, local_d)

    def compute_sets(self, full_sets=False):
        """Safe set, maximisers ``M`` and expanders ``G`` (reference: gp_opt.py:483-615).

        The three streaming passes (safe-set record, maximisers, expander candidates) are chained on the
        device: each takes the scalar it depends on (``max l0[S]``, ``max width(M)``) from the previous pass's
        records in device memory, so the host waits once, for a single 136-byte copy per rank."""
        beta = self.beta(self.t)
        if self._ci_beta is None:
            raise RuntimeError("call update_confidence_intervals() first")
        if not np.array_equal(self._ci_fmin, self.fmin):
            self.update_confidence_intervals(context=self.context)
        eng = self._engine
        G = len(self.gps)
        world = self._comm.world
        self._G_rows = []
        self._max_info = None
        self.last_trace = {}
        self._invalidate_host("S", "M")
        recs = self._record_buffers()
        m_local = self._row1 - self._row0
        tkey = (id(self.threshold), beta) if np.isscalar(self.threshold) else None
        if tkey is not None and self._thr_cache is not None and self._thr_cache[0] == tkey and self._thr_cache[1] == self.threshold:
            thr = self._thr_cache[2]
        else:
            thr = np.ascontiguousarray(np.broadcast_to(np.asarray(self.threshold, dtype=float), (G,)) * beta)
            self._thr_cache = (tkey, self.threshold, thr) if tkey is not None else None

        if not full_sets and self._cand_key_d is None:
            self._cand_key_d = eng.empty((max(m_local, 1),))
            self._cand_row_d = eng.empty((max(m_local, 1),), "i64")
        if self._fused:
            # one cooperative launch: the three passes, and between them every rank's record written straight into every
            # rank's exchange buffer over NVLink (no collective, no host round trip); one copy + one wait for all records
            host = eng.sets_fused(self._Q_d, G, self._row0, self._S_d, self.scaling, thr, not full_sets, self._M_d,
                                  None if full_sets else self._cand_key_d, None if full_sets else self._cand_row_d)
        else:
            eng.reduce_safe(self._Q_d, G, self._row0, self._S_d, self._rec_safe_l)
            self._share(self._safe_all_d, self._rec_safe_l)
            eng.maximizers_chain(self._Q_d, G, self._row0, self._S_d, self._safe_all_d, world, self.scaling, self._M_d,
                                 self._rec_max_l)
            self._share(self._max_all_d, self._rec_max_l)
            if not full_sets:
                eng.candidates_chain(self._Q_d, G, self._row0, self._S_d, self._M_d, self._max_all_d, world, self.scaling, thr,
                                     None, self._cand_key_d, self._cand_row_d, self._n_cand_l)
                self._share(self._ncand_all_d, self._n_cand_l)
            host = recs.cpu().numpy()                               # the one host wait of compute_sets
        self._check_fits()
        if world == 1:
            # one rank: nothing to combine -- unpack the two records directly (this runs once per optimize(), which is
            # host-bound on small grids)
            n_safe, max_l0, arg_l0, max_u0, arg_u0 = _SAFE_STRUCT.unpack_from(host, 0)
            self._safe_info = dict(n_safe=n_safe, max_l0=max_l0, argmax_l0=arg_l0, max_u0=max_u0, argmax_u0=arg_u0)
            if n_safe == 0:
                return
            n_max, max_w0, best_value, best_row = _MAX_STRUCT.unpack_from(host, 64)
            self._max_info = dict(n_max=n_max, max_var=max_w0 / float(self.scaling[0]), best_value=best_value, best_row=best_row)
        else:
            self._safe_info = reduce_safe_records(host[:world * 64].view(SAFE_REC_DTYPE).reshape(-1))
            if self._safe_info["n_safe"] == 0:                      # gp_opt.py:504-507 (M is already all-False: M is a subset of S)
                return
            self._max_info = reduce_max_records(host[world * 64:world * 128].view(MAX_REC_DTYPE).reshape(-1), self.scaling[0])
        max_var = self._max_info["max_var"]
        if full_sets:
            # every safe point is a candidate, natural order (gp_opt.py:527-528, :555)
            rows_local = eng.torch.nonzero(self._S_d, as_tuple=False).reshape(-1) + self._row0
            keys_local = None
            counts = self._comm.all_gather(np.array([rows_local.shape[0]], dtype=np.int64)).reshape(-1)
        else:
            counts = host[world * 128:].view(np.int64).reshape(-1)
            n_local = int(counts[self._comm.rank])
            rows_local = self._cand_row_d[:n_local]
            keys_local = self._cand_key_d[:n_local]
        n_total = int(counts.sum())
        self.last_trace = dict(max_l=self._safe_info["max_l0"], max_var=max_var, n_candidates=n_total)
        if n_total == 0:
            return
        self._expander_search(beta, rows_local, keys_local, full_sets)

    # ---- expander search ------------------------------------------------------------------
    def _ordered_candidates(self, rows_local, keys_local):
        """Global visiting order: widest (unscaled) interval first (gp_opt.py:551-552).

        Ties: the reference visits ``keys.argsort()[::-1]``.  NumPy's default argsort is not stable, so its order among
        exactly equal keys is implementation-defined; where it is defined (up to 16 candidates NumPy runs an insertion
        sort, which is stable) the reversal puts the HIGHER row first.  That is the convention used here for any size:
        key descending, then row descending (``order_candidates``)."""
        t = self._engine.torch
        if keys_local is None:
            rows = rows_local.cpu().numpy()
            allr = self._gather_var(rows)
            return np.sort(allr)
        # deterministic local order first (the append order of the candidate kernel is not): rows descending, then a
        # stable sort by key keeps the higher row first among equal keys
        order = t.argsort(rows_local, descending=True)
        rows_local, keys_local = rows_local[order], keys_local[order]
        order = t.argsort(keys_local, descending=True, stable=True)
        rows = rows_local[order].cpu().numpy()
        keys = keys_local[order].cpu().numpy()
        return order_candidates(self._comm, rows, keys)

    def _gather_var(self, arr: np.ndarray) -> np.ndarray:
        return gather_ragged(self._comm, arr)

    def _rows_values(self, rows: np.ndarray):
        """(Q rows, mean, var) of arbitrary global rows, fetched from whichever rank owns them."""
        t = self._engine.torch
        G = len(self.gps)
        rows = np.asarray(rows, dtype=np.int64)
        own = (rows >= self._row0) & (rows < self._row1)
        q = np.zeros((rows.size, 2 * G))
        mean = np.zeros((G, rows.size))
        var = np.zeros((G, rows.size))
        if own.any():
            idx = t.from_numpy(rows[own] - self._row0).to(self._engine.device)
            q[own] = self._Q_d.index_select(0, idx).cpu().numpy()
            mean[:, own] = self._mean_d.index_select(1, idx).cpu().numpy()
            var[:, own] = self._var_d.index_select(1, idx).cpu().numpy()
        if self._comm.active:
            q = self._comm.all_gather(q).sum(axis=0)
            mean = self._comm.all_gather(mean).sum(axis=0)
            var = self._comm.all_gather(var).sum(axis=0)
        return q, mean, var

    def _expander_search(self, beta, rows_local, keys_local, full_sets):
        eng = self._engine
        order = self._ordered_candidates(rows_local, keys_local)
        constrained = [i for i in range(len(self.gps)) if self.fmin[i] != -np.inf]
        m_local = self._row1 - self._row0
        B = _lib.EXPANDER_MAX_BATCH
        visited = 0
        found: List[int] = []
        for start in range(0, order.size, B):
            rows = order[start:start + B]
            nb = rows.size
            q, mean, var = self._rows_values(rows)
            xc = self._row_coordinates(rows)
            if self.num_contexts:
                xc = np.asarray(self.inputs[rows], dtype=float)
            ok = np.ones(nb, dtype=bool) if constrained else np.zeros(nb, dtype=bool)
            xc_d = eng.to_device(xc)
            for i in constrained:
                flags = eng.zeros((B,), "u8")
                if self.use_lipschitz:
                    # gp_opt.py:558-576: distance rule on the raw inputs, no GP arithmetic
                    rows_arg = None if self._grid_axes is not None else self._rows_d
                    eng.expander_lipschitz(rows_arg, xc.shape[1], self._row0, m_local, self._S_d, xc_d,
                                           eng.to_device(q[:, 2 * i + 1]), self.liptschitz[i], self.fmin[i], flags)
                    f = self._comm.any_flags(flags.cpu().numpy())[:nb].astype(bool)
                    ok &= f
                    if not ok.any() and not full_sets:
                        break
                    continue
                rows_arg = None if self._use_grid_kernel(i) else self._ensure_rows_on_device()
                if rows_arg is None:
                    self._ensure_grid_tables(i, "axes")
                eng.expander_check(i, rows_arg, self._row0, m_local, self._S_d, self._mean_d[i], self._var_d[i], xc_d,
                                   eng.to_device(mean[i]), eng.to_device(var[i]), eng.to_device(q[:, 2 * i + 1]),
                                   beta, self.fmin[i], flags)
                f = self._comm.any_flags(flags.cpu().numpy())[:nb].astype(bool)
                ok &= f
                if not ok.any() and not full_sets:
                    break
            if full_sets:
                visited += nb
                found.extend(int(r) for r in rows[ok])
            elif ok.any():
                first = int(np.flatnonzero(ok)[0])
                visited += first + 1
                found.append(int(rows[first]))      # the search stops at the first expander (gp_opt.py:611-612)
                break
            else:
                visited += nb
        self._G_rows = found
        self.last_trace.update(visited=visited, order=order)

    # ------------------------------------------------------------------ query
    def get_new_query_point(self, ucb=False):
        """Next parameters to evaluate (reference: gp_opt.py:617-649)."""
        if self._safe_info is None:
            self.compute_safe_set()
        if self._safe_info["n_safe"] == 0:
            raise EnvironmentError("There are no safe points to evaluate.")
        if ucb:
            row = self._safe_info["argmax_u0"]
        else:
            if self._max_info is None:
                raise RuntimeError("call compute_sets() before get_new_query_point()")
            value, row = self._max_info["best_value"], self._max_info["best_row"]
            if self._G_rows:
                q, _, _ = self._rows_values(np.asarray(self._G_rows, dtype=np.int64))
                gval = np.max((q[:, 1::2] - q[:, ::2]) / self.scaling, axis=1)
                vals = np.concatenate(([value], gval))
                rows = np.concatenate(([row], self._G_rows)).astype(np.int64)
                value, row = combine_max_first(vals, rows)
        self.last_query_row = int(row)
        x = self._row_point(int(row)) if not self.num_contexts else np.asarray(self.inputs[row], dtype=float)
        return x[:-self.num_contexts] if self.num_contexts else x

    def optimize(self, context=None, ucb=False):
        """One SafeOpt iteration: bounds, sets, query point (reference: gp_opt.py:651-675)."""
        self.update_confidence_intervals(context=context)
        if ucb:
            self.compute_safe_set()
        else:
            self.compute_sets()
        return self.get_new_query_point(ucb=ucb)

    def get_maximum(self, context=None):
        """Best safe lower bound and where it is, or ``None`` (reference: gp_opt.py:677-712)."""
        self.update_confidence_intervals(context=context)
        self.compute_safe_set()
        if self._safe_info["n_safe"] == 0:
            return None
        row = self._safe_info["argmax_l0"]
        x = self._row_coordinates([row])[0] if not self.num_contexts else np.asarray(self.inputs[row], dtype=float)
        return x[:-self.num_contexts or None], np.float64(self._safe_info["max_l0"])


class SafeOptSwarm(GaussianProcessOptimization):
    """SafeOpt for higher dimensions with particle swarms (reference: gp_opt.py:715-1192).

    Parameters are the reference's (``gp, fmin, bounds, beta=2, scaling='auto', threshold=0,
    swarm_size=20``) plus ``device``, ``swarm_backend`` and ``rng``.  No Lipschitz constant, no contexts
    (as in the reference).

    The particle posterior + fitness and the safe-set maintenance of ``get_new_query_point`` (the
    correlation-filtered insertion of gp_opt.py:1088-1110) run on the GPU.
    ``swarm_backend``: ``'host'`` keeps the swarm state in the reference's ``SwarmOptimization`` (NumPy
    state, fitness on the GPU; bit-for-bit the reference's random stream), ``'device'`` keeps it in HBM
    (:class:`DeviceSwarm`, sharded over the ranks of ``torch.distributed`` when initialised), ``'auto'``
    picks ``'device'`` from ``DEVICE_SWARM_MIN`` particles.  ``rng`` is passed to :class:`DeviceSwarm`.
    In a multi-rank run every rank must construct the optimiser identically and seed ``np.random``
    identically (particles are sampled from the replicated safe set with the host generator).
    """

    DEVICE_SWARM_MIN = 2048
    CORRELATION_LIMIT = 0.95        # gp_opt.py:1105

    def __init__(self, gp, fmin, bounds, beta=2, scaling="auto", threshold=0, swarm_size=20, device=None,
                 swarm_backend="auto", rng="host", seed=0, distributed=True):
        super(SafeOptSwarm, self).__init__(gp, fmin=fmin, beta=beta, num_contexts=0, threshold=threshold,
                                           scaling=scaling)
        self.S = np.asarray(self.gps[0].X)
        self.swarm_size = swarm_size
        self.max_iters = 100
        self.bounds = bounds if isinstance(bounds, list) else [bounds] * self.S.shape[1]
        self.best_lower_bound = -np.inf
        self.greedy_point = self.S[0, :]

        if swarm_backend not in ("auto", "host", "device"):
            raise ValueError("swarm_backend must be 'auto', 'host' or 'device'")
        if swarm_backend == "auto":
            swarm_backend = "device" if swarm_size >= self.DEVICE_SWARM_MIN else "host"
        self.swarm_backend = swarm_backend
        self._engine = DeviceEngine(device, max_gps=len(self.gps))
        self._comm = Comm(self._engine.device, enabled=distributed)
        self._fits = _DeviceFits(self._engine, self.gps)
        self._fit_buffers = {}
        self._fit_buffers_keep = {}
        self.optimal_velocities = self.optimize_particle_velocity()
        swarm_types = ["greedy", "maximizers", "expanders"]
        if swarm_backend == "device":
            self._peer = bool(hasattr(self._engine, "connect_exchange") and self._engine.connect_exchange(self._comm))
            self.swarms = {kind: DeviceSwarm(self._engine, self.optimal_velocities, partial(self._swarm_fitness, kind),
                                             bounds=self.bounds, rng=rng, seed=seed + 1000 * k, comm=self._comm, peer=self._peer)
                           for k, kind in enumerate(swarm_types)}
            for sw in self.swarms.values():
                sw.fitness_key = self._fitness_key
        else:
            self.swarms = {kind: SwarmOptimization(swarm_size, self.optimal_velocities,
                                                   partial(self._compute_particle_fitness, kind), bounds=self.bounds)
                           for kind in swarm_types}

    def optimize_particle_velocity(self):
        """Per-dimension velocities at which the prior correlation drops to ~0.95 (gp_opt.py:818-872).

        O(G*d*30) scalar kernel evaluations, host side like the reference."""
        d = self.gp.input_dim
        origin = np.zeros((1, d), dtype=float)
        vel = np.empty((len(self.gps), d), dtype=float)
        for i, gp in enumerate(self.gps):
            for j in range(d):
                probe = np.zeros((1, d), dtype=float)
                hi, lo = 1000.0, 0.0
                while True:
                    mid = (hi + lo) / 2
                    probe[0, j] = mid
                    corr = np.asarray(gp.kern.K(origin, probe)).squeeze() / self.scaling[i] ** 2
                    enough = corr > 0.94
                    slow = corr < 0.95
                    if slow:
                        hi = mid
                    elif enough:
                        lo = mid
                    if (slow and enough) or hi - lo < 1e-5:
                        break
                vel[i, j] = mid
        out = np.min(vel, axis=0)
        out /= np.sqrt(d)
        return out

    def _compute_penalty(self, slack):
        """Constraint-violation penalty (gp_opt.py:874-899); host helper kept for API parity."""
        slack = np.atleast_1d(np.asarray(slack, dtype=float))
        pen = np.clip(slack, None, 0)
        pen[(slack < 0) & (slack > -0.001)] *= 2
        pen[(slack <= -0.001) & (slack > -0.1)] *= 5
        pen[(slack <= -0.1) & (slack > -1)] *= 10
        big = slack < -1
        pen[big] = -300 * pen[big] ** 2
        return pen

    def _fitness_buffers(self, P, keep=False):
        """Per-swarm-size scratch (posterior planes, values, flags), reused across the ~300 fitness passes of an optimize().
        ``keep``: the buffers of the device swarms are never evicted -- a captured PSO iteration (CUDA graph) holds their
        addresses, and an evicted buffer would be handed to the next tensor of that size while the graph still writes to it."""
        buf = self._fit_buffers_keep.get(P) or self._fit_buffers.get(P)
        if buf is None:
            eng, G = self._engine, len(self.gps)
            buf = (eng.empty((G, P)), eng.empty((G, P)), eng.empty((P,)), eng.empty((P,), "u8"))
            if keep:
                self._fit_buffers_keep[P] = buf
            else:
                if len(self._fit_buffers) > 4:
                    self._fit_buffers.clear()
                self._fit_buffers[P] = buf
        elif keep and P not in self._fit_buffers_keep:
            self._fit_buffers_keep[P] = buf
        return buf

    def _fitness_device(self, swarm_type, particles_d, fresh=True, keep=False):
        """Fitness of device-resident particles; returns device tensors (values, safe).  The returned tensors are scratch
        that the next call with the same particle count overwrites (``DeviceSwarm`` consumes them at once).  Inside a swarm
        run the GP objects cannot change, so ``fresh=False`` skips the fingerprint check of the device fits."""
        eng = self._engine
        if fresh:
            self._fits.refresh()
        beta = self.beta(self.t)
        P = particles_d.shape[0]
        G = len(self.gps)
        mean, var, values, safe = self._fitness_buffers(P, keep=keep)
        n_needed = 1 if swarm_type == "greedy" else G
        for group in self._fits.groups:
            group = [i for i in group if i < n_needed]
            if len(group) > 1 and eng.posterior_multi(group, particles_d, 0, P, beta, [-np.inf] * len(group),
                                                      means=[mean[i] for i in group], variances=[var[i] for i in group]):
                continue
            for i in group:
                eng.posterior_rows(i, particles_d, beta, -np.inf, mean=mean[i], var=var[i])
        eng.swarm_fitness(_lib.SWARM_KINDS[swarm_type], G if n_needed == G else 1, P, mean, var, beta,
                          self.fmin[:n_needed] if n_needed == G else self.fmin[:1], self.scaling[:max(n_needed, 1)],
                          self.best_lower_bound, values, safe)
        return values, safe

    def _fitness_key(self):
        """Everything ``_fitness_device`` passes to its launches by value: a captured PSO iteration is valid while this is
        unchanged (a new observation, new thresholds or a new greedy bound re-capture it)."""
        f = self._fits
        return (float(self.beta(self.t)), float(self.best_lower_bound), self.fmin.tobytes(), np.asarray(self.scaling).tobytes(),
                f.refits, f.appends, f.removals, tuple(tuple(g) for g in f.groups), self.t)

    def _swarm_fitness(self, swarm_type, particles_d):
        """Fitness callback of the device swarms: the fits were refreshed by get_new_query_point before the run."""
        return self._fitness_device(swarm_type, particles_d, fresh=False, keep=True)

    def _compute_particle_fitness(self, swarm_type, particles):
        """Fitness value and safety flag of every particle (reference: gp_opt.py:901-1013).

        ``swarm_type`` is 'greedy', 'maximizers', 'expanders' or 'safe_set'."""
        if swarm_type not in _lib.SWARM_KINDS:
            raise AssertionError("Invalid swarm type")
        particles = np.ascontiguousarray(np.atleast_2d(np.asarray(particles, dtype=float)))
        values, safe = self._fitness_device(swarm_type, self._engine.to_device(particles))
        # fresh host arrays: the device tensors are scratch that the next fitness pass overwrites, and SwarmOptimization keeps
        # the returned values as its best_values (swarm.py:80)
        return np.array(values.cpu().numpy(), copy=True), safe.cpu().numpy().astype(bool)

    def get_new_query_point(self, swarm_type):
        """Run one swarm and return (point, value / std-devs) (reference: gp_opt.py:1015-1134)."""
        beta = self.beta(self.t)
        safe_size, input_dim = self.S.shape

        _, safe = self._compute_particle_fitness("safe_set", self.S)
        num_safe = safe.sum()
        if num_safe == 0:
            raise RuntimeError("The safe set is empty.")
        if num_safe >= self.swarm_size and num_safe != len(safe):
            logging.warning("Warning: {} unsafe points removed. Model might be violated"
                            .format(np.count_nonzero(~safe)))
            self.S = self.S[safe]
            safe_size = self.S.shape[0]

        if swarm_type == "greedy":
            random_id = np.random.randint(safe_size, size=self.swarm_size - 3)
            best_sampled_point = np.argmax(self.gp.Y)
            particles = np.vstack((self.S[random_id, :], self.greedy_point, self.gp.X[-1, :],
                                   self.gp.X[best_sampled_point]))
        else:
            random_id = np.random.randint(safe_size, size=self.swarm_size)
            particles = self.S[random_id, :]

        swarm = self.swarms[swarm_type]
        swarm.init_swarm(particles)
        swarm.run_swarm(self.max_iters)
        on_device = isinstance(swarm, DeviceSwarm)
        global_best = swarm.global_best                      # host array in both backends
        best_value = swarm.global_best_value if on_device else np.max(swarm.best_values)

        if swarm_type != "greedy":
            new_rows = self._select_new_safe_points(swarm.best_positions, sharded=on_device)
            if new_rows.shape[0]:
                self.S = np.vstack((self.S, new_rows))
            logging.debug("At the end of swarm {}, {} points were appended to the safeset".format(swarm_type,
                                                                                                   new_rows.shape[0]))
        else:
            mean, var = self._posterior_host(0, self.greedy_point[None, :])
            lower_bound = mean.squeeze() - beta * np.sqrt(var.squeeze())
            if lower_bound < best_value:
                self.greedy_point = global_best.copy()

        if swarm_type == "greedy":
            return global_best.copy(), best_value

        var = np.empty(len(self.gps), dtype=float)
        for i in range(len(self.gps)):
            var[i] = self._posterior_host(i, global_best[None, :])[1].squeeze()
        return global_best, np.sqrt(var)

    def _select_new_safe_points(self, best_positions, sharded=False):
        """Rows of ``best_positions`` that join the safe set (reference: gp_opt.py:1088-1110).

        The reference materialises ``K(best_positions, vstack(S, best_positions)) / scaling[0]**2`` and walks
        the particles in index order, accepting one iff its prior correlation with every point already in the
        set is ``<= 0.95``.  Here: a tiled filter against the old safe set (each rank on its own particles when
        the swarm is sharded), one all-gather of (position, keep) over the ranks, then the blocked sequential
        walk over the survivors (``so_safeset_insert``, run redundantly on every rank so that ``S`` stays
        replicated and identical).  Returns the accepted positions as a host array, in particle order."""
        eng = self._engine
        t = eng.torch
        self._fits.refresh()
        cand = best_positions if t.is_tensor(best_positions) else eng.to_device(np.ascontiguousarray(best_positions, dtype=float))
        n_local, d = cand.shape
        scale2 = float(self.scaling[0]) ** 2
        keep = eng.empty((max(n_local, 1),), "u8")[:n_local]
        eng.safeset_filter(0, cand, eng.to_device(np.ascontiguousarray(self.S, dtype=float)), scale2,
                           self.CORRELATION_LIMIT, keep)
        if sharded and self._comm.active:
            cand = gather_padded_rows(self._comm, cand, self.swarm_size)
            keep = gather_padded_rows(self._comm, keep, self.swarm_size)
        n = cand.shape[0]
        accept, acc_pos, n_acc = eng.empty((max(n, 1),), "u8"), eng.empty((max(n, 1), d)), eng.zeros((1,), "i64")
        eng.safeset_insert(0, cand, keep, scale2, self.CORRELATION_LIMIT, accept, acc_pos, n_acc)
        count = int(n_acc.item())
        return acc_pos[:count].cpu().numpy()

    def _posterior_host(self, i, X):
        """Posterior of GP ``i`` at a few host points through the device path."""
        eng = self._engine
        self._fits.refresh()
        Xd = eng.to_device(np.ascontiguousarray(np.asarray(X, dtype=float)))
        mean, var = eng.empty((Xd.shape[0],)), eng.empty((Xd.shape[0],))
        eng.posterior_rows(i, Xd, 0.0, -np.inf, mean=mean, var=var)
        return mean.cpu().numpy(), var.cpu().numpy()

    def optimize(self, ucb=False):
        """One SafeOptSwarm iteration (reference: gp_opt.py:1136-1177)."""
        self.greedy, self.best_lower_bound = self.get_new_query_point("greedy")
        x_maxi, std_maxi = self.get_new_query_point("maximizers")
        if ucb:
            logging.info("Using ucb criterion.")
            return x_maxi
        x_exp, std_exp = self.get_new_query_point("expanders")
        std_exp[(std_exp < self.threshold) | (self.fmin == -np.inf)] = 0
        std_exp /= self.scaling
        std_exp = np.max(std_exp)
        std_maxi = std_maxi[0] / self.scaling[0]
        logging.info("The best maximizer has std. dev. %f" % std_maxi)
        logging.info("The best expander has std. dev. %f" % std_exp)
        logging.info("The greedy estimate of lower bound has value %f" % self.best_lower_bound)
        return x_maxi if std_maxi > std_exp else x_exp

    def get_maximum(self):
        """Best observed point (reference: gp_opt.py:1179-1192)."""
        best = np.argmax(self.gp.Y)
        return self.gp.X[best, :], self.gp.Y[best]
