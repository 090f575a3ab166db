"""Grid construction and recognition.

``linearly_spaced_combinations`` has the reference's signature and row order
(/root/reference/safeopt/utilities.py:21-54).  ``detect_grid`` recognises a parameter set that
is a Cartesian product in that row order, so the device can regenerate rows from the row index
(no M x d array in HBM, separable-kernel tables for RBF) -- the check is bitwise, and a
parameter set that fails it simply takes the explicit-rows path.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import List, Optional

import numpy as np

__all__ = ["linearly_spaced_combinations", "detect_grid", "grid_row_strides", "grid_rows_from_index"]


def linearly_spaced_combinations(bounds, num_samples):
    """All combinations of per-dimension ``np.linspace`` samples, one combination per row.

    Parameters
    ----------
    bounds : sequence of (min, max) pairs, one per variable.
    num_samples : int or sequence of ints -- samples per variable.

    Returns
    -------
    (prod(num_samples), len(bounds)) array.  For d >= 2 the rows follow ``np.meshgrid``'s default
    'xy' indexing: variable 1 varies slowest, then variable 0, then variables 2..d-1 (fastest).
    """
    d = len(bounds)
    if not isinstance(num_samples, Sequence):
        num_samples = [num_samples] * d
    axes = [np.linspace(b[0], b[1], int(n)) for b, n in zip(bounds, num_samples)]
    if d == 1:
        return axes[0][:, None]
    return np.array([g.ravel() for g in np.meshgrid(*axes)]).T


def grid_row_strides(n: Sequence[int]) -> List[int]:
    """stride_j such that row = sum_j idx_j * stride_j in the reference row order."""
    d = len(n)
    if d == 1:
        return [1]
    strides = [0] * d
    s = 1
    for j in range(d - 1, 1, -1):
        strides[j] = s
        s *= int(n[j])
    strides[0] = s
    s *= int(n[0])
    strides[1] = s
    return strides


def grid_rows_from_index(axes: Sequence[np.ndarray], rows) -> np.ndarray:
    """Coordinates of the given (global) row indices of the product grid ``axes``."""
    rows = np.atleast_1d(np.asarray(rows, dtype=np.int64))
    n = [len(a) for a in axes]
    strides = grid_row_strides(n)
    out = np.empty((rows.size, len(axes)))
    for j, a in enumerate(axes):
        out[:, j] = np.asarray(a)[(rows // strides[j]) % n[j]]
    return out


def detect_grid(parameter_set: np.ndarray, max_dim: int = 6) -> Optional[List[np.ndarray]]:
    """Return the per-axis values if ``parameter_set`` is bit-identical to their product grid
    in reference row order, else ``None``."""
    ps = np.asarray(parameter_set)
    if ps.ndim != 2 or ps.dtype != np.float64:
        return None
    M, d = ps.shape
    if d < 1 or d > max_dim or M < 1:
        return None
    # read the axis values off the rows where only that index moves (cheap), then verify everything
    if d == 1:
        axes = [ps[:, 0].copy()]
        return axes if np.all(np.diff(axes[0]) != 0) or M == 1 else None
    first = ps[0]
    # fastest axis is d-1 (or axis 0 when d == 2): count its length as the run before it wraps
    order = [1, 0] + list(range(2, d))          # slowest ... fastest
    n = [0] * d
    stride = 1
    for j in reversed(order):
        col = ps[::stride, j] if stride > 1 else ps[:, j]
        # length of axis j = index of the first return to its first value
        same = np.flatnonzero(col[1:] == col[0])
        nj = int(same[0]) + 1 if same.size else col.size
        n[j] = nj
        stride *= nj
    if stride != M:
        return None
    strides = grid_row_strides(n)
    axes = [ps[0:strides[j] * n[j]:strides[j], j].copy() for j in range(d)]
    if any(len(np.unique(a)) != len(a) for a in axes):
        return None
    # bitwise verification, chunked to bound memory
    idx = np.arange(M, dtype=np.int64)
    for j in range(d):
        if not np.array_equal(axes[j][(idx // strides[j]) % n[j]], ps[:, j]):
            return None
    del first
    return axes
