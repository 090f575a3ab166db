"""Seeded synthetic problems for the benchmark and parity tests.

Pure data: NumPy arrays and hyper-parameters, no GP arithmetic.  The five named
configurations are BASELINE.json's ``configs``:

  C1  d=1, 100-point grid, N=5,   G=1           (examples/1d_example.ipynb shape; CPU-runnable)
  C2  d=2, 200x200 grid,   N=64,  G=1, fp64
  C3  d=2, 500x500 grid,   N=128, G=3, fp32
  C4  d=4, 50^4 grid,      N=256, G=1, fp64     (the headline; rows sharded across GPUs)
  C5  swarm d=6, 1e5 particles, N=512, G=2
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


@dataclass
class GridWorkload:
    name: str
    d: int
    num_samples: object       # per axis: int, or a list with one entry per axis
    n_train: int
    n_gps: int
    dtype: str
    bounds: List[Tuple[float, float]]
    X: np.ndarray             # (N, d)
    Y: np.ndarray             # (N, G)
    variance: float
    lengthscale: np.ndarray   # (d,) ARD
    noise_var: float
    fmin: List[float]
    beta: float
    threshold: float

    @property
    def n_rows(self) -> int:
        return int(np.prod(self.samples_per_axis))

    @property
    def samples_per_axis(self) -> List[int]:
        if isinstance(self.num_samples, (list, tuple)):
            return [int(n) for n in self.num_samples]
        return [int(self.num_samples)] * self.d


def _objective(X: np.ndarray) -> np.ndarray:
    return 2.0 * np.exp(-np.sum(X * X, axis=1) / 8.0)


def grid_workload(name: str, d: int, num_samples: int, n_train: int, n_gps: int = 1,
                  dtype: str = "fp64", seed: int = 0) -> GridWorkload:
    """RBF-ARD problem of SURVEY.md section 8d: bounds [-5,5]^d, sigma_f^2 = 2, l = 1,
    noise 0.05^2, X ~ U(-1.5, 1.5)^d, f(x) = 2 exp(-|x|^2/8), Y_i = f(X) + 0.05 randn(seed+i)."""
    rs = np.random.RandomState(seed)
    X = rs.uniform(-1.5, 1.5, size=(n_train, d))
    cols = []
    for i in range(n_gps):
        noise = np.random.RandomState(seed + i).randn(n_train) if i else rs.randn(n_train)
        cols.append(_objective(X) + 0.05 * noise)
    Y = np.stack(cols, axis=1)
    return GridWorkload(name=name, d=d, num_samples=num_samples, n_train=n_train, n_gps=n_gps, dtype=dtype,
                        bounds=[(-5.0, 5.0)] * d, X=X, Y=Y, variance=2.0, lengthscale=np.ones(d),
                        noise_var=0.05 ** 2, fmin=[0.0] * n_gps, beta=2.0, threshold=0.2)


def config(name: str, seed: int = 0, num_samples=None) -> GridWorkload:
    table = {
        "C1": dict(d=1, num_samples=100, n_train=5, n_gps=1, dtype="fp64"),
        "C2": dict(d=2, num_samples=200, n_train=64, n_gps=1, dtype="fp64"),
        "C3": dict(d=2, num_samples=500, n_train=128, n_gps=3, dtype="fp32"),
        "C4": dict(d=4, num_samples=50, n_train=256, n_gps=1, dtype="fp64"),
    }
    kw = dict(table[name])
    if num_samples is not None:
        kw["num_samples"] = num_samples
    return grid_workload(name, seed=seed, **kw)


@dataclass
class SwarmWorkload:
    name: str
    d: int
    n_particles: int
    n_train: int
    n_gps: int
    bounds: List[Tuple[float, float]]
    X: np.ndarray
    Y: np.ndarray
    particles: np.ndarray
    variance: float
    lengthscale: np.ndarray
    noise_var: float
    fmin: List[float]
    beta: float


def swarm_workload(n_particles: int = 100_000, n_train: int = 512, d: int = 6, n_gps: int = 2,
                   seed: int = 0) -> SwarmWorkload:
    """C5: particles ~ U(-1,1)^6, X ~ U(-0.5,0.5)^6, f = 1 - 0.3 |x|^2."""
    rs = np.random.RandomState(seed)
    X = rs.uniform(-0.5, 0.5, size=(n_train, d))
    f = 1.0 - 0.3 * np.sum(X * X, axis=1)
    Y = np.stack([f + 0.05 * np.random.RandomState(seed + 1 + i).randn(n_train) for i in range(n_gps)], axis=1)
    particles = rs.uniform(-1.0, 1.0, size=(n_particles, d))
    return SwarmWorkload(name="C5", d=d, n_particles=n_particles, n_train=n_train, n_gps=n_gps,
                         bounds=[(-1.0, 1.0)] * d, X=X, Y=Y, particles=particles, variance=2.0,
                         lengthscale=np.ones(d), noise_var=0.05 ** 2, fmin=[0.0] * n_gps, beta=2.0)
