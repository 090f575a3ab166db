"""safeopt_b200 -- B200-native implementation of SafeOpt's per-iteration hot path.

Same public names as the reference package (/root/reference/safeopt/__init__.py:36-39):
``SafeOpt``, ``SafeOptSwarm``, ``linearly_spaced_combinations`` -- plus the swarm driver and a
minimal GPy-protocol model (``GPRegression``, ``RBF``, ``Matern32``, ``Matern52``) for
environments without GPy.  Importing the package does not need a GPU; constructing an optimiser
does, and fails loudly without one (there is no CPU execution path).
"""
from __future__ import absolute_import

from .utilities import linearly_spaced_combinations  # noqa: F401
from .gpmodel import GPRegression, RBF, Matern32, Matern52, ProductKernel, kern, models  # noqa: F401
from .swarm import SwarmOptimization, DeviceSwarm  # noqa: F401
from .gp_opt import SafeOpt, SafeOptSwarm, GaussianProcessOptimization  # noqa: F401
from ._lib import NativeLibraryError, DeviceError  # noqa: F401

__all__ = ["SafeOpt", "SafeOptSwarm", "GaussianProcessOptimization", "linearly_spaced_combinations",
           "SwarmOptimization", "DeviceSwarm", "GPRegression", "RBF", "Matern32", "Matern52", "ProductKernel"]
__version__ = "0.1.0"
