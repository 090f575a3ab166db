"""Import the UNMODIFIED reference package from /root/reference on Python 3.12 / NumPy 2.

TEST INFRASTRUCTURE ONLY (oracle).  Nothing here is copied from the reference: the
reference's own modules are imported from where they lie, after three compatibility
shims that its age requires:

  * ``collections.Sequence`` (removed in Python 3.10; used at safeopt/utilities.py:9 and
    safeopt/gp_opt.py:10),
  * ``np.float`` (removed in NumPy 1.24; used at safeopt/gp_opt.py:376, 829-836, 957, 967,
    1129 and safeopt/swarm.py:54,58),
  * stub ``matplotlib`` / ``mpl_toolkits`` modules (plotting only, never called here),

and ``GPy`` is provided by :mod:`oracle.gpy_lite` (real GPy is not installable offline).

``/root/reference`` exists only in the build container, never on the GPU box, so this
module is used solely by ``oracle/make_golden.py`` and by tests that skip when the
reference tree is absent.
"""
from __future__ import annotations

import collections
import collections.abc
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SAFEOPT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "safeopt", "gp_opt.py"))


def _stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__["__getattr__"] = lambda attr: _Anything()
    return mod


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, item):
        return _Anything()


def import_reference():
    """Return the reference ``safeopt`` package (imported from REFERENCE_ROOT)."""
    if not reference_available():
        raise ImportError("reference tree not found at %s" % REFERENCE_ROOT)
    import numpy as np

    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence
    if "float" not in np.__dict__:
        np.float = float  # noqa: NPY001 - deliberate shim
    if "bool" not in np.__dict__:
        np.bool = bool
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = _stub(name)
    from . import gpy_lite

    sys.modules.setdefault("GPy", gpy_lite)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("safeopt")
