"""GP model objects speaking the GPy protocol SafeOpt relies on, and the adapter that reads
hyper-parameters out of any such object (these classes, real GPy models, or test doubles).

The reference never does GP arithmetic itself; it duck-types on a GPy model
(/root/reference/safeopt/gp_opt.py:58, :347, :780; protocol listed in SURVEY.md section 8b):
``.X .Y .input_dim .kern.input_dim .kern.K .kern.Kdiag .set_XY .predict_noiseless``.
GPy is not installable here, so this module ships a minimal model with the same surface whose
``predict_noiseless`` runs on the GPU through the C ABI (so_fit + so_posterior_rows).  The
``kern.K`` / ``kern.Kdiag`` helpers are host NumPy: the reference only calls them on a handful
of points (gp_opt.py:83 scaling, :847 velocity bisection, :1093 swarm de-duplication) and they
are not part of the device hot path.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional

import numpy as np

from . import _lib

_KIND_BY_NAME = {
    "rbf": _lib.KERNEL_RBF, "RBF": _lib.KERNEL_RBF,
    "Mat32": _lib.KERNEL_MATERN32, "Matern32": _lib.KERNEL_MATERN32,
    "Mat52": _lib.KERNEL_MATERN52, "Matern52": _lib.KERNEL_MATERN52,
}


class UnsupportedModelError(TypeError):
    """The GP object uses a kernel / likelihood the device path does not implement."""


# --------------------------------------------------------------------------- kernels
class StationaryKernel:
    """Stationary ARD kernel: k(x, x') = variance * f(|| (x - x') / lengthscale ||)."""

    kind = None
    name = "stationary"

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, active_dims=None):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        if lengthscale is None:
            lengthscale = 1.0
        ls = np.atleast_1d(np.asarray(lengthscale, dtype=float)).ravel()
        if self.ARD and ls.size == 1:
            ls = np.full(self.input_dim, ls[0])
        if not self.ARD and ls.size != 1:
            raise ValueError("a non-ARD kernel takes one lengthscale")
        self.lengthscale = ls
        self.variance = np.atleast_1d(np.asarray(variance, dtype=float))
        self.active_dims = np.arange(self.input_dim) if active_dims is None else np.asarray(active_dims, dtype=int)

    def _profile(self, r2):
        raise NotImplementedError

    def _r2(self, X, X2):
        ls = self.lengthscale if self.ARD else np.full(self.active_dims.size, self.lengthscale[0])
        A = np.asarray(X, dtype=float)[:, self.active_dims] / ls
        B = A if X2 is None else np.asarray(X2, dtype=float)[:, self.active_dims] / ls
        diff = A[:, None, :] - B[None, :, :]
        return np.einsum("ijk,ijk->ij", diff, diff)

    def K(self, X, X2=None):
        """Host-side covariance for small inputs (see module docstring)."""
        return float(self.variance[0]) * self._profile(self._r2(X, X2))

    def Kdiag(self, X):
        return np.full(np.asarray(X).shape[0], float(self.variance[0]))

    def __mul__(self, other):
        return ProductKernel([self, other])


class RBF(StationaryKernel):
    kind = _lib.KERNEL_RBF
    name = "rbf"

    def _profile(self, r2):
        return np.exp(-0.5 * r2)


class Matern32(StationaryKernel):
    kind = _lib.KERNEL_MATERN32
    name = "Mat32"

    def _profile(self, r2):
        r = np.sqrt(r2)
        return (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)


class Matern52(StationaryKernel):
    kind = _lib.KERNEL_MATERN52
    name = "Mat52"

    def _profile(self, r2):
        r = np.sqrt(r2)
        return (1.0 + np.sqrt(5.0) * r + 5.0 / 3.0 * r2) * np.exp(-np.sqrt(5.0) * r)


class ProductKernel:
    """Product of kernels on disjoint ``active_dims`` (contexts, examples/context_example.ipynb)."""

    name = "mul"

    def __init__(self, parts):
        flat: List[StationaryKernel] = []
        for p in parts:
            flat.extend(p.parts if isinstance(p, ProductKernel) else [p])
        self.parts = flat
        self.input_dim = int(max(int(np.max(p.active_dims)) for p in flat) + 1)
        self.active_dims = np.arange(self.input_dim)

    def K(self, X, X2=None):
        out = 1.0
        for p in self.parts:
            out = out * p.K(X, X2)
        return out

    def Kdiag(self, X):
        out = 1.0
        for p in self.parts:
            out = out * p.Kdiag(X)
        return out

    def __mul__(self, other):
        return ProductKernel([self, other])


class _Likelihood:
    def __init__(self, variance):
        self.variance = np.atleast_1d(np.asarray(variance, dtype=float))


# --------------------------------------------------------------------------- hyper-parameter adapter
class Hyper(NamedTuple):
    kind: int
    lengthscale: np.ndarray      # (d,) ARD-expanded
    variance: float
    noise_var: float


def _scalar(x) -> float:
    return float(np.asarray(x, dtype=float).reshape(-1)[0])


def _kind_of(k) -> Optional[int]:
    kind = getattr(k, "kind", None)
    if isinstance(kind, int):
        return kind
    for key in (type(k).__name__, getattr(k, "name", None)):
        if key in _KIND_BY_NAME:
            return _KIND_BY_NAME[key]
    return None


def _stationary_part(k, d_total):
    """(kind, lengthscale over d_total dims or nan where inactive, variance)."""
    kind = _kind_of(k)
    if kind is None:
        raise UnsupportedModelError(
            "kernel %r is not supported by the device path (RBF, Matern32, Matern52 and products of RBFs only; "
            "there is no CPU fallback)" % (type(k).__name__,))
    dims = np.asarray(getattr(k, "active_dims", np.arange(int(k.input_dim))), dtype=int).ravel()
    ls = np.asarray(k.lengthscale, dtype=float).ravel()
    if ls.size == 1:
        ls = np.full(dims.size, ls[0])
    if ls.size != dims.size:
        raise UnsupportedModelError("lengthscale size %d does not match active_dims %d" % (ls.size, dims.size))
    full = np.full(d_total, np.nan)
    full[dims] = ls
    return kind, full, _scalar(k.variance)


def extract_hyper(gp) -> Hyper:
    """Read kernel family, ARD lengthscales, signal and noise variance from a GPy-like model.

    Attribute names follow GPy 1.x (``kern.variance``, ``kern.lengthscale``, ``kern.ARD``,
    ``kern.active_dims``, ``kern.parts`` for products, ``likelihood.variance``).  Anything else
    raises :class:`UnsupportedModelError` -- loudly, because there is no CPU path to fall back to."""
    d = int(np.asarray(gp.X).shape[1])
    k = gp.kern
    try:
        noise = _scalar(gp.likelihood.variance)
    except AttributeError as exc:
        raise UnsupportedModelError("model has no Gaussian likelihood.variance") from exc
    # GPy's GP applies `mean_function` and `normalizer` inside predict_noiseless; the device posterior has neither, so a
    # model that uses them would silently get different bounds.  Refuse it (GPy's normalizer=False/None means "off").
    if getattr(gp, "mean_function", None) is not None:
        raise UnsupportedModelError("models with a mean_function are not supported by the device path")
    if getattr(gp, "normalizer", None) not in (None, False):
        raise UnsupportedModelError("models with an output normalizer are not supported by the device path")
    parts = getattr(k, "parts", None)
    if parts is None or _kind_of(k) is not None:
        kind, ls, var = _stationary_part(k, d)
        if np.isnan(ls).any():
            raise UnsupportedModelError("kernel does not cover all %d input dimensions" % d)
        return Hyper(kind, ls, var, noise)
    # composite kernel: only a PRODUCT is separable (GPy: class Prod, name 'mul'); a sum (class Add, name 'sum') or any
    # other combination has a different covariance and must not be folded into one ARD kernel
    if not (type(k).__name__ in ("Prod", "ProductKernel") or getattr(k, "name", None) == "mul"):
        raise UnsupportedModelError(
            "composite kernel %r (name %r) is not supported by the device path: only products of RBF kernels on disjoint "
            "active_dims are (there is no CPU fallback)" % (type(k).__name__, getattr(k, "name", None)))
    # product kernel: RBF x RBF on disjoint dims == one ARD RBF with the variances multiplied
    ls = np.full(d, np.nan)
    var = 1.0
    for part in parts:
        kind, pls, pvar = _stationary_part(part, d)
        if kind != _lib.KERNEL_RBF:
            raise UnsupportedModelError("only products of RBF kernels are separable into one ARD kernel")
        mask = ~np.isnan(pls)
        if np.any(~np.isnan(ls[mask])):
            raise UnsupportedModelError("product kernel parts overlap in active_dims")
        ls[mask] = pls[mask]
        var *= pvar
    if np.isnan(ls).any():
        raise UnsupportedModelError("product kernel does not cover all %d input dimensions" % d)
    return Hyper(_lib.KERNEL_RBF, ls, var, noise)


def fingerprint(gp, hyper: Hyper):
    """Cheap identity of the training-side state, to decide whether the device fit is stale."""
    X = np.ascontiguousarray(np.asarray(gp.X, dtype=float))
    Y = np.ascontiguousarray(np.asarray(gp.Y, dtype=float))
    return (X.shape, hash(X.tobytes()), hash(Y.tobytes()), hyper.kind, hyper.lengthscale.tobytes(), hyper.variance,
            hyper.noise_var)


# --------------------------------------------------------------------------- model
class GPRegression:
    """Exact GP regression, GPy ``GPRegression`` surface, posterior evaluated on the GPU."""

    def __init__(self, X, Y, kernel=None, noise_var=1.0, device=None):
        X = np.atleast_2d(np.asarray(X, dtype=float))
        Y = np.atleast_2d(np.asarray(Y, dtype=float))
        self.kern = RBF(X.shape[1]) if kernel is None else kernel
        self.likelihood = _Likelihood(noise_var)
        self.input_dim = X.shape[1]
        self._device = device
        self._engine = None
        self._fp = None
        self.set_XY(X, Y)

    def set_XY(self, X, Y):
        """Replace the data (gp_opt.py:227,:267,:275); the device refit happens lazily on next use."""
        self.X = np.array(X, dtype=float, copy=True)
        self.Y = np.array(Y, dtype=float, copy=True)

    def _ensure_fit(self):
        from .engine import DeviceEngine
        if self._engine is None:
            self._engine = DeviceEngine(self._device, max_gps=1)
        hyper = extract_hyper(self)
        fp = fingerprint(self, hyper)
        if fp != self._fp:
            self._engine.fit(0, self.X, self.Y[:, 0], hyper.kind, hyper.lengthscale, hyper.variance, hyper.noise_var)
            self._fp = fp
        return self._engine

    def predict_noiseless(self, Xnew, full_cov=False):
        """Posterior mean and variance, shapes (M,1),(M,1) like GPy (reference call site gp_opt.py:469)."""
        if full_cov:
            raise NotImplementedError("full_cov is not part of the SafeOpt hot path")
        eng = self._ensure_fit()
        Xnew = np.ascontiguousarray(np.atleast_2d(np.asarray(Xnew, dtype=float)))
        Xd = eng.to_device(Xnew)
        M = Xnew.shape[0]
        mean, var = eng.empty((M,)), eng.empty((M,))
        eng.posterior_rows(0, Xd, 0.0, -np.inf, mean=mean, var=var)
        return mean.cpu().numpy()[:, None], var.cpu().numpy()[:, None]

    def _raw_predict(self, Xnew, full_cov=False):
        return self.predict_noiseless(Xnew, full_cov=full_cov)


class _Namespace:
    pass


kern = _Namespace()
kern.RBF, kern.Matern32, kern.Matern52, kern.Prod = RBF, Matern32, Matern52, ProductKernel
models = _Namespace()
models.GPRegression = GPRegression
