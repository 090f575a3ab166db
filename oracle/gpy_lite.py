"""CPU restatement of the slice of GPy that SafeOpt calls.  TEST INFRASTRUCTURE ONLY.

This file is part of the *oracle*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``safeopt_b200``) never does.

Parity status: **parity unpinned at the GPy level** -- GPy (``GPy>=0.8``, unpinned in
/root/reference/requirements.txt:1, not vendored) is absent from this environment and the
reference's own tests hold no posterior golden vectors.  The
restatement below follows the published GPy 1.x algorithm; it is cross-checked in
``tests/test_oracle.py`` against closed forms (N=1) and against scikit-learn's independent
``GaussianProcessRegressor``.

Call sites in the reference that define the surface restated here
(all in /root/reference/safeopt/gp_opt.py): ``kern.Kdiag`` :83, ``set_XY`` :227 :267 :275,
``predict_noiseless`` :469 :591 :929 :973 :1117 :1132, ``kern.K`` :847 :1093,
``.X/.Y`` :121-126, ``.input_dim`` :82, ``kern.input_dim`` :152.

GPy semantics restated (GPy 1.x):
  * stationary kernels use the expansion-form distance
    r^2 = |a|^2 + |b|^2 - 2 a.b, clipped at 0, diagonal forced to 0 for K(X, X);
    ARD divides inputs by the length-scale vector first, otherwise r is divided afterwards.
  * exact inference adds ``noise + 1e-8`` to the diagonal, factors with LAPACK dpotrf,
    alpha = dpotrs(L, Y), and the predictive variance uses the *explicit* inverse from
    dpotri (symmetrised):  var = Kdiag - colsum((W^T Kx) * Kx), clipped to >= 1e-15.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack

JITTER = 1e-8          # always added to the noise variance by GPy's exact inference
VAR_FLOOR = 1e-15      # GPy clips predictive variances here


# --------------------------------------------------------------------------- kernels
class _Param(np.ndarray):
    """Tiny stand-in for paramz.Param: an ndarray that also behaves like a float when size 1."""

    def __new__(cls, value):
        return np.atleast_1d(np.asarray(value, dtype=float)).view(cls)

    def __float__(self):
        return float(np.asarray(self).reshape(-1)[0])


class Stationary:
    """Base for kernels that are functions of the scaled distance r only."""

    name = "stationary"

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, active_dims=None):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        if lengthscale is None:
            lengthscale = np.ones(self.input_dim if self.ARD else 1)
        ls = np.atleast_1d(np.asarray(lengthscale, dtype=float))
        if self.ARD and ls.size == 1:
            ls = np.full(self.input_dim, float(ls[0]))
        if not self.ARD and ls.size != 1:
            raise ValueError("non-ARD kernel takes exactly one lengthscale")
        self.lengthscale = _Param(ls)
        self.variance = _Param(variance)
        if active_dims is None:
            active_dims = np.arange(self.input_dim)
        self.active_dims = np.asarray(active_dims, dtype=int)

    # -- distance ------------------------------------------------------------------
    def _slice(self, X):
        X = np.asarray(X, dtype=float)
        if X.shape[1] == self.active_dims.size and np.array_equal(self.active_dims, np.arange(X.shape[1])):
            return X
        return X[:, self.active_dims]

    @staticmethod
    def _unscaled_dist(A, B=None):
        if B is None:
            sq = np.sum(np.square(A), 1)
            r2 = -2.0 * A.dot(A.T) + (sq[:, None] + sq[None, :])
            np.fill_diagonal(r2, 0.0)
        else:
            asq = np.sum(np.square(A), 1)
            bsq = np.sum(np.square(B), 1)
            r2 = -2.0 * A.dot(B.T) + (asq[:, None] + bsq[None, :])
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)

    def _scaled_dist(self, A, B=None):
        if self.ARD:
            ls = np.asarray(self.lengthscale)
            return self._unscaled_dist(A / ls, None if B is None else B / ls)
        return self._unscaled_dist(A, B) / float(self.lengthscale)

    # -- public GPy surface ----------------------------------------------------------
    def K(self, X, X2=None):
        A = self._slice(X)
        B = None if X2 is None else self._slice(X2)
        return self.K_of_r(self._scaled_dist(A, B))

    def Kdiag(self, X):
        out = np.empty(np.asarray(X).shape[0])
        out[:] = float(self.variance)
        return out

    def K_of_r(self, r):  # pragma: no cover - abstract
        raise NotImplementedError

    def __mul__(self, other):
        return Prod([self, other])


class RBF(Stationary):
    name = "rbf"

    def K_of_r(self, r):
        return float(self.variance) * np.exp(-0.5 * r ** 2)


class Matern32(Stationary):
    name = "Mat32"

    def K_of_r(self, r):
        s3 = np.sqrt(3.0)
        return float(self.variance) * (1.0 + s3 * r) * np.exp(-s3 * r)


class Matern52(Stationary):
    name = "Mat52"

    def K_of_r(self, r):
        s5 = np.sqrt(5.0)
        return float(self.variance) * (1.0 + s5 * r + 5.0 / 3.0 * r ** 2) * np.exp(-s5 * r)


class Prod:
    """Elementwise product of kernels, each on its own ``active_dims`` (context example)."""

    name = "mul"

    def __init__(self, parts):
        flat = []
        for p in parts:
            flat.extend(p.parts if isinstance(p, Prod) else [p])
        self.parts = flat
        self.input_dim = int(max(int(p.active_dims.max()) for p in flat) + 1)
        self.active_dims = np.arange(self.input_dim)

    def K(self, X, X2=None):
        out = None
        for p in self.parts:
            k = p.K(X, X2)
            out = k if out is None else out * k
        return out

    def Kdiag(self, X):
        out = None
        for p in self.parts:
            k = p.Kdiag(X)
            out = k if out is None else out * k
        return out

    def __mul__(self, other):
        return Prod([self, other])


class _Kern:
    """Namespace mirroring ``GPy.kern``."""
    RBF = RBF
    Matern32 = Matern32
    Matern52 = Matern52
    Prod = Prod


kern = _Kern


# --------------------------------------------------------------------------- likelihood
class Gaussian:
    name = "Gaussian_noise"

    def __init__(self, variance=1.0):
        self.variance = _Param(variance)


# --------------------------------------------------------------------------- linear algebra
def jitchol(A, maxtries=5):
    """Lower Cholesky with GPy's jitter-retry ladder."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        L, info = lapack.dpotrf(A + np.eye(A.shape[0]) * jitter, lower=1)
        if info == 0:
            return L
        jitter *= 10
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def inverse_from_chol(L):
    """Explicit symmetric inverse from the lower factor via dpotri (GPy: dpotri + symmetrify)."""
    Ai, info = lapack.dpotri(np.asfortranarray(L), lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri failed")
    Ai = np.tril(Ai)
    return Ai + np.tril(Ai, -1).T


# --------------------------------------------------------------------------- model
class GPRegression:
    """Exact GP regression with a Gaussian likelihood (GPy.models.GPRegression surface)."""

    def __init__(self, X, Y, kernel=None, noise_var=1.0):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        Y = np.atleast_2d(np.asarray(Y, dtype=float))
        if kernel is None:
            kernel = RBF(X.shape[1])
        self.kern = kernel
        self.likelihood = Gaussian(noise_var)
        self.input_dim = X.shape[1]
        self.set_XY(X, Y)

    # reference call sites gp_opt.py:227,267,275 -- a full refit every time
    def set_XY(self, X, Y):
        self.X = np.array(X, dtype=float, copy=True)
        self.Y = np.array(Y, dtype=float, copy=True)
        self._refit()

    def _refit(self):
        Ky = self.kern.K(self.X).copy()
        Ky[np.diag_indices_from(Ky)] += float(self.likelihood.variance) + JITTER
        self._L = jitchol(Ky)
        alpha, info = lapack.dpotrs(self._L, self.Y, lower=1)
        if info != 0:
            raise np.linalg.LinAlgError("dpotrs failed")
        self._alpha = alpha
        self._Winv = None

    @property
    def woodbury_inv(self):
        if self._Winv is None:
            self._Winv = inverse_from_chol(self._L)
        return self._Winv

    @property
    def woodbury_vector(self):
        return self._alpha

    @property
    def woodbury_chol(self):
        return self._L

    def _raw_predict(self, Xnew, full_cov=False):
        Xnew = np.asarray(Xnew, dtype=float)
        Kx = self.kern.K(self.X, Xnew)
        mu = Kx.T.dot(self._alpha)
        if mu.ndim == 1:
            mu = mu.reshape(-1, 1)
        if full_cov:
            Kxx = self.kern.K(Xnew)
            var = Kxx - Kx.T.dot(self.woodbury_inv.dot(Kx))
            return mu, var
        Kxx = self.kern.Kdiag(Xnew)
        var = (Kxx - np.sum(self.woodbury_inv.T.dot(Kx) * Kx, 0))[:, None]
        var = np.clip(var, VAR_FLOOR, np.inf)
        return mu, var

    def predict_noiseless(self, Xnew, full_cov=False):
        return self._raw_predict(Xnew, full_cov=full_cov)

    def predict(self, Xnew, full_cov=False, include_likelihood=True):
        mu, var = self._raw_predict(Xnew, full_cov=full_cov)
        if include_likelihood:
            var = var + float(self.likelihood.variance)
        return mu, var


class _Models:
    GPRegression = GPRegression


models = _Models


def posterior_chunked(gp, Xnew, chunk=250_000):
    """``predict_noiseless`` in row chunks (bounded memory; results unchanged row by row)."""
    M = Xnew.shape[0]
    mean = np.empty((M, 1))
    var = np.empty((M, 1))
    for s in range(0, M, chunk):
        m, v = gp.predict_noiseless(Xnew[s:s + chunk])
        mean[s:s + chunk] = m
        var[s:s + chunk] = v
    return mean, var
