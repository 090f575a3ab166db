"""ctypes binding of the C ABI declared in include/safeopt_b200.h.

The shared library is built in-tree by :mod:`safeopt_b200.build` (``csrc/libsafeopt_b200.so``).
There is no CPU fallback: if the library is missing or fails to load, importing anything that
needs the device raises :class:`NativeLibraryError`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SAFEOPT_B200_LIB points at an alternative build of the same ABI (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("SAFEOPT_B200_LIB") or os.path.join(HERE, "csrc", "libsafeopt_b200.so")

SO_OK = 0
SO_ERR_BAD_ARG = -1
SO_ERR_UNSUPPORTED = -2
SO_ERR_NOT_PD = -3
SO_ERR_CUDA = -4
SO_ERR_NOT_FITTED = -5
SO_ERR_CAPACITY = -6
SO_ERR_NO_DEVICE = -7
SO_ERR_TIMEOUT = -8

KERNEL_RBF, KERNEL_MATERN32, KERNEL_MATERN52 = 0, 1, 2
SAFE_NONE, SAFE_WRITE, SAFE_AND = 0, 1, 2
SWARM_GREEDY, SWARM_MAXIMIZERS, SWARM_EXPANDERS, SWARM_SAFE_SET = 0, 1, 2, 3
SWARM_KINDS = {"greedy": SWARM_GREEDY, "maximizers": SWARM_MAXIMIZERS, "expanders": SWARM_EXPANDERS,
               "safe_set": SWARM_SAFE_SET}
EXPANDER_MAX_BATCH = 32
ABI_VERSION = 4
SWARM_REC_DOUBLES = 18
XCHG_HANDLE_BYTES = 64
XCHG_MAX_WORLD = 16


def sets_result_bytes(world: int) -> int:
    """SO_SETS_RESULT_BYTES of the header."""
    return world * 136 + 16


class NativeLibraryError(RuntimeError):
    """The sm_100a extension is missing or unusable (the product has no CPU path)."""


class DeviceError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, status: int, where: str, detail: str):
        self.status = status
        super().__init__("%s failed with status %d (%s)%s" % (where, status, status_string(status),
                                                               (": " + detail) if detail else ""))


class SafeRecord(C.Structure):
    _fields_ = [("n_safe", C.c_int64), ("max_l0", C.c_double), ("argmax_l0", C.c_int64),
                ("max_u0", C.c_double), ("argmax_u0", C.c_int64), ("reserved", C.c_int64 * 3)]


class MaxRecord(C.Structure):
    _fields_ = [("n_max", C.c_int64), ("max_width0", C.c_double), ("best_value", C.c_double),
                ("best_row", C.c_int64), ("reserved", C.c_int64 * 4)]


_P = C.c_void_p
_i, _i64, _dbl, _u64 = C.c_int, C.c_int64, C.c_double, C.c_uint64

# name -> (restype, argtypes); exactly the declarations of include/safeopt_b200.h
SIGNATURES = {
    "so_abi_version": (_i, []),
    "so_status_string": (C.c_char_p, [_i]),
    "so_create": (_i, [_i, _i, C.POINTER(_P)]),
    "so_destroy": (_i, [_P]),
    "so_last_error": (C.c_char_p, [_P]),
    "so_num_sms": (_i, [_P]),
    "so_fit": (_i, [_P, _i, _P, _P, _i, _i, _i, _P, _dbl, _dbl, _P]),
    "so_fit_async": (_i, [_P, _i, _P, _P, _i, _i, _i, _P, _dbl, _dbl, _P]),
    "so_fit_status": (_i, [_P, _i]),
    "so_fit_like": (_i, [_P, _i, _i, _P, _P]),
    "so_fit_append": (_i, [_P, _i, _P, _dbl, _P]),
    "so_fit_remove_last": (_i, [_P, _i, _P]),
    "so_fit_export": (_i, [_P, _i, _P, _P, _P]),
    "so_grid_define": (_i, [_P, _i, _P, _P, _P]),
    "so_grid_prepare": (_i, [_P, _i, _P]),
    "so_grid_prepare_rows": (_i, [_P, _i, _i64, _i64, _P]),
    "so_posterior_rows": (_i, [_P, _i, _P, _i64, _dbl, _dbl, _P, _P, _P, _i, _i, _P, _i, _P]),
    "so_posterior_grid": (_i, [_P, _i, _i64, _i64, _dbl, _dbl, _P, _P, _P, _i, _i, _P, _i, _P]),
    "so_posterior_rows_multi": (_i, [_P, _i, _P, _P, _i64, _dbl, _P, _P, _P, _P, _i, _P, _P, _i, _P]),
    "so_posterior_grid_multi": (_i, [_P, _i, _P, _i64, _i64, _dbl, _P, _P, _P, _P, _i, _P, _P, _i, _P]),
    "so_grid_prepare_f32": (_i, [_P, _i, _i64, _i64, _P]),
    "so_posterior_grid_f32": (_i, [_P, _i, _P, _i64, _i64, _dbl, _P, _P, _P, _P, _i, _P, _P, _i, _P]),
    "so_debug_row_plan": (_i, [_i, _P, _P]),
    "so_debug_row_plan_slots": (_i, [_i, _i, _P, _P]),
    "so_debug_tile_plans": (_i, [_i, _i, _i64, _i, _i64, _i, _P]),
    "so_posterior_rows_simple": (_i, [_P, _i, _P, _i64, _P, _P, _P]),
    "so_grid_rows": (_i, [_P, _i64, _i64, _P, _P]),
    "so_sets_reduce_safe": (_i, [_P, _P, _i, _i64, _i64, _P, _P, _P]),
    "so_sets_maximizers": (_i, [_P, _P, _i, _i64, _i64, _P, _dbl, _P, _P, _P, _P]),
    "so_sets_candidates": (_i, [_P, _P, _i, _i64, _i64, _P, _P, _dbl, _P, _P, _P, _P, _P, _i64, _P, _P]),
    "so_sets_maximizers_chain": (_i, [_P, _P, _i, _i64, _i64, _P, _P, _i, _P, _P, _P, _P]),
    "so_sets_candidates_chain": (_i, [_P, _P, _i, _i64, _i64, _P, _P, _P, _i, _P, _P, _P, _P, _P, _i64, _P, _P]),
    "so_sets_fused": (_i, [_P, _P, _i, _i64, _i64, _P, _P, _P, _i, _P, _P, _P, _i64, _P, _P]),
    "so_sets_fused_result": (_i, [_P, _P, _P]),
    "so_debug_fused_times": (_i, [_P, _P]),
    "so_xchg_export": (_i, [_P, _P]),
    "so_xchg_connect": (_i, [_P, _i, _i, _P]),
    "so_xchg_world": (_i, [_P]),
    "so_expander_check": (_i, [_P, _i, _P, _i64, _i64, _P, _P, _P, _P, _P, _P, _P, _i, _dbl, _dbl, _P, _P]),
    "so_expander_lipschitz": (_i, [_P, _P, _i, _i64, _i64, _P, _P, _P, _i, _dbl, _dbl, _P, _P]),
    "so_swarm_fitness": (_i, [_P, _i, _i, _i64, _P, _P, _dbl, _P, _P, _dbl, _P, _P, _P]),
    "so_swarm_step": (_i, [_P, _i64, _i, _P, _P, _P, _P, _P, _dbl, _P, _P, _P]),
    "so_swarm_update_best": (_i, [_P, _i64, _i, _P, _P, _P, _P, _P, _P, _i64, _P, _P]),
    "so_swarm_combine_best": (_i, [_P, _P, _i, _i, _P, _P, _P]),
    "so_swarm_rand": (_i, [_P, _i64, _i, _i64, _u64, _u64, _P, _P]),
    "so_swarm_step_dev": (_i, [_P, _i64, _i, _i64, _P, _P, _P, _P, _P, _u64, _P, _P, _P]),
    "so_swarm_update_best_x": (_i, [_P, _i64, _i, _P, _P, _P, _P, _P, _P, _i64, _P, _P, _P, _P]),
    "so_safeset_filter": (_i, [_P, _i, _P, _i64, _P, _i64, _dbl, _dbl, _P, _P]),
    "so_safeset_insert": (_i, [_P, _i, _P, _i64, _P, _dbl, _dbl, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library with argtypes set."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            "CUDA extension not built: %s is missing. Run `python -m safeopt_b200.build` "
            "(there is no CPU fallback)." % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the machine
        raise NativeLibraryError("cannot load %s: %s" % (LIB_PATH, exc)) from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError("%s does not export %s" % (LIB_PATH, name)) from exc
        fn.restype = res
        fn.argtypes = args
    if lib.so_abi_version() != ABI_VERSION:
        raise NativeLibraryError("ABI version mismatch: library %d, binding %d" % (lib.so_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def status_string(status: int) -> str:
    return load().so_status_string(status).decode()
