"""Device engine: one C-ABI handle plus the torch tensors that own every device buffer.

PyTorch is plumbing here (allocation, streams, host<->device copies, torch.distributed); every
arithmetic step of the hot path is a call into ``csrc/libsafeopt_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import DeviceError, NativeLibraryError

SAFE_REC_DTYPE = np.dtype([("n_safe", "<i8"), ("max_l0", "<f8"), ("argmax_l0", "<i8"), ("max_u0", "<f8"),
                           ("argmax_u0", "<i8"), ("reserved", "<i8", (3,))])
MAX_REC_DTYPE = np.dtype([("n_max", "<i8"), ("max_width0", "<f8"), ("best_value", "<f8"), ("best_row", "<i8"),
                          ("reserved", "<i8", (4,))])
assert SAFE_REC_DTYPE.itemsize == 64 and MAX_REC_DTYPE.itemsize == 64


def _np_f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def _hptr(a: Optional[np.ndarray]) -> C.c_void_p:
    return C.c_void_p(0 if a is None else a.ctypes.data)


class DeviceEngine:
    """Owns a ``so_handle`` on one CUDA device."""

    def __init__(self, device=None, max_gps: int = 8):
        try:
            import torch
        except ImportError as exc:  # pragma: no cover
            raise NativeLibraryError("PyTorch is required for device memory management") from exc
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise NativeLibraryError("no CUDA device visible: safeopt_b200 has no CPU execution path")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if device.type != "cuda":
            raise NativeLibraryError("safeopt_b200 runs on CUDA devices only (got %s)" % device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.max_gps = int(max_gps)
        handle = C.c_void_p()
        rc = self.lib.so_create(device.index, self.max_gps, C.byref(handle))
        if rc != 0:
            raise DeviceError(rc, "so_create", "")
        self.handle = handle
        self.num_sms = self.lib.so_num_sms(handle)
        self.launches = 0          # kernels launched through this engine (bench bookkeeping)
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        self._grid_axes = None
        self._tape = None
        self._pending_fits = set()
        self.xchg_world = 1
        self._sets_result = np.zeros(_lib.sets_result_bytes(1), dtype=np.uint8)

    # ------------------------------------------------------------------ helpers
    def close(self):
        if getattr(self, "handle", None):
            self.lib.so_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> C.c_void_p:
        """Raw handle of torch's current stream on this device (every kernel is launched on it).  The private raw getter is
        ~30x cheaper than building a torch.cuda.Stream object per launch, which matters for the launch-bound loops (a PSO
        iteration issues seven kernels)."""
        if self._raw_stream is not None:
            return C.c_void_p(self._raw_stream(self.device.index))
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc: int, where: str):
        if rc != 0:
            detail = self.lib.so_last_error(self.handle).decode(errors="replace")
            raise DeviceError(rc, where, detail)

    def empty(self, shape, dtype="f64"):
        t = self.torch
        dt = {"f64": t.float64, "u8": t.uint8, "i64": t.int64, "f32": t.float32}[dtype]
        return t.empty(shape, dtype=dt, device=self.device)

    def zeros(self, shape, dtype="f64"):
        t = self.torch
        dt = {"f64": t.float64, "u8": t.uint8, "i64": t.int64, "f32": t.float32}[dtype]
        return t.zeros(shape, dtype=dt, device=self.device)

    def to_device(self, arr: np.ndarray, pinned: bool = False):
        t = self.torch.from_numpy(np.ascontiguousarray(arr))
        if pinned:
            t = t.pin_memory()
        return t.to(self.device, non_blocking=pinned)

    def synchronize(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ K1
    def fit(self, gp: int, X, Y, kind: int, lengthscale, variance: float, noise_var: float, wait: bool = True):
        """``wait=False``: do not wait for the factorisation (so_fit_async); a failed Cholesky is then reported by
        :meth:`check_fits`, which the caller runs after its next synchronisation."""
        X = _np_f64(X)
        Y = _np_f64(Y).reshape(-1)
        N, d = X.shape
        ls = _np_f64(lengthscale).reshape(-1)
        if ls.size == 1:
            ls = np.full(d, float(ls[0]))
        if ls.size != d:
            raise ValueError("lengthscale must have 1 or %d entries" % d)
        fn = self.lib.so_fit if wait else self.lib.so_fit_async
        rc = fn(self.handle, gp, _hptr(X), _hptr(Y), N, d, kind, _hptr(ls), float(variance), float(noise_var), self._stream())
        self._check(rc, "so_fit")
        if not wait:
            self._pending_fits.add(gp)
        panels = (8 * ((N + 7) // 8) + 31) // 32           # fit.cu: k_chol_panel per panel, k_chol_update between panels
        self.launches += 4 + 2 * panels - 1

    def check_fits(self):
        """Status of the fits started with ``wait=False`` (call after a synchronisation of the stream)."""
        pending, self._pending_fits = self._pending_fits, set()
        for gp in pending:
            self._check(self.lib.so_fit_status(self.handle, gp), "so_fit")

    def fit_like(self, gp: int, src_gp: int, Y):
        """Fit of a GP that shares inputs, kernel and noise with the fitted ``src_gp``: copies its factorisation."""
        Y = _np_f64(Y).reshape(-1)
        self._check(self.lib.so_fit_like(self.handle, gp, src_gp, _hptr(Y), self._stream()), "so_fit_like")
        self.launches += 1

    def fit_append(self, gp: int, x_new, y_new: float) -> bool:
        """One-point update of an existing fit (f4).  Returns False when the device asks for a full refit
        (buffers full, or the bordered matrix lost positive definiteness)."""
        x = _np_f64(x_new).reshape(-1)
        rc = self.lib.so_fit_append(self.handle, gp, _hptr(x), float(y_new), self._stream())
        if rc in (_lib.SO_ERR_CAPACITY, _lib.SO_ERR_NOT_PD):
            return False
        self._check(rc, "so_fit_append")
        self.launches += 4
        return True

    def fit_remove_last(self, gp: int):
        self._check(self.lib.so_fit_remove_last(self.handle, gp, self._stream()), "so_fit_remove_last")
        self.launches += 3

    def fit_export(self, gp: int, N: int):
        L = np.empty((N, N))
        Linv = np.empty((N, N))
        alpha = np.empty(N)
        self._check(self.lib.so_fit_export(self.handle, gp, _hptr(L), _hptr(Linv), _hptr(alpha)), "so_fit_export")
        return L, Linv, alpha

    # ------------------------------------------------------------------ grid
    def define_grid(self, axes: Sequence[np.ndarray]):
        n = np.asarray([len(a) for a in axes], dtype=np.int32)
        vals = _np_f64(np.concatenate([np.asarray(a, dtype=np.float64) for a in axes]))
        self._check(self.lib.so_grid_define(self.handle, len(axes), _hptr(n), _hptr(vals), self._stream()), "so_grid_define")
        self._grid_axes = [np.asarray(a, dtype=np.float64) for a in axes]

    def prepare_grid(self, gp: int, row0: Optional[int] = None, n_rows: Optional[int] = None):
        """Grid tables of GP ``gp``; with a row range, the per-slow-index operands only for that row block."""
        if row0 is None:
            self._check(self.lib.so_grid_prepare(self.handle, gp, self._stream()), "so_grid_prepare")
        else:
            self._check(self.lib.so_grid_prepare_rows(self.handle, gp, int(row0), int(n_rows), self._stream()),
                        "so_grid_prepare_rows")
        self.launches += 4

    def grid_rows(self, row0: int, M: int):
        d = len(self._grid_axes)
        out = self.empty((M, d))
        self._check(self.lib.so_grid_rows(self.handle, row0, M, _ptr(out), self._stream()), "so_grid_rows")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ K2
    # A "tape" remembers the posterior launches of one update_confidence_intervals() with their converted arguments, so that
    # the next call with unchanged inputs re-issues them without rebuilding anything: on small grids (and on eight GPUs, where a
    # rank's K2 takes 1.7 ms) the host time BEFORE the K2 launch is GPU idle time, because the previous step ended with a host wait.
    def start_tape(self):
        self._tape = []

    def stop_tape(self):
        tape, self._tape = self._tape, None
        return tape

    def replay_tape(self, tape):
        st = self._stream()
        for fn, args, _keep, where, n in tape:
            rc = fn(*args, st)
            if rc:
                self._check(rc, where)
            self.launches += n

    def _k2(self, fn, where, args, keep, n_launch):
        rc = fn(*args, self._stream())
        if rc == 0 and self._tape is not None:
            self._tape.append((fn, args, keep, where, n_launch))
        return rc

    def posterior_rows(self, gp, Xstar, beta, fmin, mean=None, var=None, Q=None, q_col=0, S=None, safe_mode=_lib.SAFE_NONE):
        M = Xstar.shape[0]
        q_stride = 0 if Q is None else Q.shape[1]
        rc = self._k2(self.lib.so_posterior_rows, "so_posterior_rows",
                      (self.handle, gp, _ptr(Xstar), M, float(beta), float(fmin), _ptr(mean), _ptr(var), _ptr(Q), q_stride, q_col,
                       _ptr(S), safe_mode), (Xstar, mean, var, Q, S), 1 if M else 0)
        self._check(rc, "so_posterior_rows")
        self.launches += 1 if M else 0

    def posterior_grid(self, gp, row0, M, beta, fmin, mean=None, var=None, Q=None, q_col=0, S=None, safe_mode=_lib.SAFE_NONE):
        q_stride = 0 if Q is None else Q.shape[1]
        rc = self._k2(self.lib.so_posterior_grid, "so_posterior_grid",
                      (self.handle, gp, int(row0), int(M), float(beta), float(fmin), _ptr(mean), _ptr(var), _ptr(Q), q_stride, q_col,
                       _ptr(S), safe_mode), (mean, var, Q, S), 1 if M else 0)
        self._check(rc, "so_posterior_grid")
        self.launches += 1 if M else 0

    def posterior_multi(self, gps, Xstar, row0, M, beta, fmins, means=None, variances=None, Q=None, q_cols=None, S=None,
                        safe_mode=_lib.SAFE_NONE) -> bool:
        """Several GPs that share X, kernel and noise in ONE launch (one contraction, one V.z per GP).  ``Xstar`` = explicit
        rows or ``None`` for the defined grid.  Returns False when the device has no room for the extra outputs (the caller
        then evaluates the GPs one by one)."""
        n = len(gps)
        gi = (C.c_int * n)(*[int(g) for g in gps])
        fm = _np_f64(fmins)
        qc = (C.c_int * n)(*[int(c) for c in (q_cols if q_cols is not None else [0] * n)])
        mp = (C.c_void_p * n)(*[0 if means is None or t is None else t.data_ptr() for t in (means or [None] * n)])
        vp = (C.c_void_p * n)(*[0 if variances is None or t is None else t.data_ptr() for t in (variances or [None] * n)])
        q_stride = 0 if Q is None else Q.shape[1]
        keep = (gi, fm, qc, mp, vp, Xstar, means, variances, Q, S)
        if Xstar is None:
            rc = self._k2(self.lib.so_posterior_grid_multi, "so_posterior_multi",
                          (self.handle, n, gi, int(row0), int(M), float(beta), _hptr(fm), mp, vp, _ptr(Q), q_stride, qc, _ptr(S), safe_mode),
                          keep, 1 if M else 0)
        else:
            rc = self._k2(self.lib.so_posterior_rows_multi, "so_posterior_multi",
                          (self.handle, n, gi, _ptr(Xstar), int(M), float(beta), _hptr(fm), mp, vp, _ptr(Q), q_stride, qc, _ptr(S), safe_mode),
                          keep, 1 if M else 0)
        if rc == _lib.SO_ERR_CAPACITY:
            return False
        self._check(rc, "so_posterior_multi")
        self.launches += 1 if M else 0
        return True

    def prepare_grid_f32(self, gp: int, row0: int, n_rows: int):
        """TF32 hi/lo operand planes of the fp32 tensor-core grid kernel for GP ``gp`` and this row block."""
        self._check(self.lib.so_grid_prepare_f32(self.handle, gp, int(row0), int(n_rows), self._stream()), "so_grid_prepare_f32")
        self.launches += 3

    def posterior_grid_f32(self, gps, row0, M, beta, fmins, means=None, variances=None, Q=None, q_cols=None, S=None,
                           safe_mode=_lib.SAFE_NONE):
        """posterior_multi on the defined grid in fp32 arithmetic (tcgen05 3xTF32, TMEM accumulators); one GP or several that
        share the factorisation."""
        n = len(gps)
        gi = (C.c_int * n)(*[int(g) for g in gps])
        fm = _np_f64(fmins)
        qc = (C.c_int * n)(*[int(c) for c in (q_cols if q_cols is not None else [0] * n)])
        mp = (C.c_void_p * n)(*[0 if means is None or t is None else t.data_ptr() for t in (means or [None] * n)])
        vp = (C.c_void_p * n)(*[0 if variances is None or t is None else t.data_ptr() for t in (variances or [None] * n)])
        q_stride = 0 if Q is None else Q.shape[1]
        rc = self._k2(self.lib.so_posterior_grid_f32, "so_posterior_grid_f32",
                      (self.handle, n, gi, int(row0), int(M), float(beta), _hptr(fm), mp, vp, _ptr(Q), q_stride, qc, _ptr(S), safe_mode),
                      (gi, fm, qc, mp, vp, means, variances, Q, S), 2 if M else 0)
        self._check(rc, "so_posterior_grid_f32")
        self.launches += 2 if M else 0            # k_mean_grid (fp64 means) + k_posterior_f32

    def posterior_rows_simple(self, gp, Xstar):
        M = Xstar.shape[0]
        mean, var = self.empty((M,)), self.empty((M,))
        self._check(self.lib.so_posterior_rows_simple(self.handle, gp, _ptr(Xstar), M, _ptr(mean), _ptr(var), self._stream()),
                    "so_posterior_rows_simple")
        return mean, var

    # ------------------------------------------------------------------ K3
    def reduce_safe(self, Q, n_gps, row0, S, rec):
        self._check(self.lib.so_sets_reduce_safe(self.handle, _ptr(Q), n_gps, Q.shape[0], int(row0), _ptr(S), _ptr(rec),
                                                 self._stream()), "so_sets_reduce_safe")
        self.launches += 1

    def maximizers(self, Q, n_gps, row0, S, max_l0, scaling, Mmask, rec):
        sc = _np_f64(scaling)
        self._check(self.lib.so_sets_maximizers(self.handle, _ptr(Q), n_gps, Q.shape[0], int(row0), _ptr(S), float(max_l0),
                                                _hptr(sc), _ptr(Mmask), _ptr(rec), self._stream()), "so_sets_maximizers")
        self.launches += 1

    def candidates(self, Q, n_gps, row0, S, Mmask, max_var, scaling, thr, cand_mask, cand_key, cand_row, n_cand):
        sc, th = _np_f64(scaling), _np_f64(thr)
        cap = 0 if cand_key is None else cand_key.shape[0]
        self._check(self.lib.so_sets_candidates(self.handle, _ptr(Q), n_gps, Q.shape[0], int(row0), _ptr(S), _ptr(Mmask),
                                                float(max_var), _hptr(sc), _hptr(th), _ptr(cand_mask), _ptr(cand_key),
                                                _ptr(cand_row), cap, _ptr(n_cand), self._stream()), "so_sets_candidates")
        self.launches += 1

    def maximizers_chain(self, Q, n_gps, row0, S, safe_recs, n_recs, scaling, Mmask, rec):
        sc = _np_f64(scaling)
        self._check(self.lib.so_sets_maximizers_chain(self.handle, _ptr(Q), n_gps, Q.shape[0], int(row0), _ptr(S),
                                                      _ptr(safe_recs), int(n_recs), _hptr(sc), _ptr(Mmask), _ptr(rec),
                                                      self._stream()), "so_sets_maximizers_chain")
        self.launches += 1

    def candidates_chain(self, Q, n_gps, row0, S, Mmask, max_recs, n_recs, scaling, thr, cand_mask, cand_key, cand_row, n_cand):
        sc, th = _np_f64(scaling), _np_f64(thr)
        cap = 0 if cand_key is None else cand_key.shape[0]
        self._check(self.lib.so_sets_candidates_chain(self.handle, _ptr(Q), n_gps, Q.shape[0], int(row0), _ptr(S), _ptr(Mmask),
                                                      _ptr(max_recs), int(n_recs), _hptr(sc), _hptr(th), _ptr(cand_mask),
                                                      _ptr(cand_key), _ptr(cand_row), cap, _ptr(n_cand), self._stream()),
                    "so_sets_candidates_chain")
        self.launches += 1

    def sets_fused(self, Q, n_gps, row0, S, scaling, thr, with_candidates, Mmask, cand_key, cand_row, fetch=True):
        """The three set passes and their cross-rank record exchange in one launch (so_sets_fused).  With ``fetch`` the
        call waits for the kernel and returns the combined records of all ranks as a uint8 array laid out
        [world x safe record][world x max record][world x int64 count]; otherwise use :meth:`sets_fused_result`."""
        sc, th = _np_f64(scaling), _np_f64(thr)
        cap = 0 if cand_key is None else cand_key.shape[0]
        out = self._sets_result if fetch else None
        rc = self.lib.so_sets_fused(self.handle, _ptr(Q), int(n_gps), Q.shape[0], int(row0), _ptr(S), _hptr(sc), _hptr(th),
                                    1 if with_candidates else 0, _ptr(Mmask), _ptr(cand_key), _ptr(cand_row), cap,
                                    _hptr(out), self._stream())
        self._check(rc, "so_sets_fused")
        self.launches += 1
        return out[:136 * self.xchg_world] if fetch else None

    def sets_fused_result(self):
        out = self._sets_result
        self._check(self.lib.so_sets_fused_result(self.handle, _hptr(out), self._stream()), "so_sets_fused_result")
        return out[:136 * self.xchg_world]

    # ------------------------------------------------------------------ cross-rank exchange
    def connect_exchange(self, comm) -> bool:
        """Map the peers' exchange buffers (CUDA IPC) so that the fused kernels exchange their records over NVLink
        themselves.  Collective over ``comm`` (every rank constructs the same objects in the same order).  Returns False --
        after saying so once on stderr -- when IPC mapping is unavailable; the callers then keep the NCCL all-gather path."""
        if not comm.active:
            return True
        import os
        import sys
        if os.environ.get("SAFEOPT_B200_PEER_EXCHANGE", "1") == "0":
            return False
        mine = np.zeros(_lib.XCHG_HANDLE_BYTES, dtype=np.uint8)
        rc = self.lib.so_xchg_export(self.handle, _hptr(mine))
        ok = np.array([1 if rc == 0 else 0], dtype=np.int64)
        handles = comm.all_gather(mine)                           # (world, 64) uint8, rank order
        if comm.all_gather(ok).min() == 1:
            rc = self.lib.so_xchg_connect(self.handle, comm.world, comm.rank, _hptr(np.ascontiguousarray(handles)))
            ok[0] = 1 if rc == 0 else 0
        good = bool(comm.all_gather(ok).min() == 1)               # all ranks take the same path
        if not good:
            if rc == 0:                                           # a peer failed: fall back to a single-rank view here too
                self.lib.so_xchg_connect(self.handle, 1, 0, _hptr(mine))
            if comm.rank == 0:
                sys.stderr.write("safeopt_b200: CUDA IPC peer mapping unavailable (%s); cross-rank records go through "
                                 "torch.distributed all-gathers instead\n" % self.lib.so_last_error(self.handle).decode(errors="replace"))
            return False
        self.xchg_world = comm.world
        self._sets_result = np.zeros(_lib.sets_result_bytes(comm.world), dtype=np.uint8)
        comm.barrier()                                            # nobody launches a fused kernel before everyone is mapped
        return True

    # ------------------------------------------------------------------ K4
    def expander_check(self, gp, Xstar, row0, M, S, mean, var, xc, mean_c, var_c, u_c, beta, fmin, flags):
        B = xc.shape[0]
        rc = self.lib.so_expander_check(self.handle, gp, _ptr(Xstar), int(row0), int(M), _ptr(S), _ptr(mean), _ptr(var),
                                        _ptr(xc), _ptr(mean_c), _ptr(var_c), _ptr(u_c), B, float(beta), float(fmin),
                                        _ptr(flags), self._stream())
        self._check(rc, "so_expander_check")
        self.launches += 2 if M else 0

    def expander_lipschitz(self, Xstar, d, row0, M, S, xc, u_c, lipschitz, fmin, flags):
        rc = self.lib.so_expander_lipschitz(self.handle, _ptr(Xstar), int(d), int(row0), int(M), _ptr(S), _ptr(xc), _ptr(u_c),
                                            xc.shape[0], float(lipschitz), float(fmin), _ptr(flags), self._stream())
        self._check(rc, "so_expander_lipschitz")
        self.launches += 1 if M else 0

    # ------------------------------------------------------------------ K5/K6
    def swarm_fitness(self, kind, n_gps, P, mean, var, beta, fmin, scaling, best_lower_bound, values, safe):
        fm, sc = _np_f64(fmin), _np_f64(scaling)
        rc = self.lib.so_swarm_fitness(self.handle, kind, n_gps, int(P), _ptr(mean), _ptr(var), float(beta), _hptr(fm),
                                       _hptr(sc), float(best_lower_bound), _ptr(values), _ptr(safe), self._stream())
        self._check(rc, "so_swarm_fitness")
        self.launches += 1

    def swarm_step(self, pos, vel, best_pos, global_best, r, inertia, velocity_scale, bounds):
        P, d = pos.shape
        vs = _np_f64(velocity_scale)
        bd = None if bounds is None else _np_f64(bounds)
        rc = self.lib.so_swarm_step(self.handle, P, d, _ptr(pos), _ptr(vel), _ptr(best_pos), _ptr(global_best), _ptr(r),
                                    float(inertia), _hptr(vs), _hptr(bd), self._stream())
        self._check(rc, "so_swarm_step")
        self.launches += 1

    def swarm_update_best(self, pos, values, safe, best_pos, best_values, best_idx, p0=0, rec=None):
        P, d = pos.shape
        rc = self.lib.so_swarm_update_best(self.handle, P, d, _ptr(pos), _ptr(values), _ptr(safe), _ptr(best_pos),
                                           _ptr(best_values), _ptr(best_idx), int(p0), _ptr(rec), self._stream())
        self._check(rc, "so_swarm_update_best")
        self.launches += 1

    def swarm_rand(self, P, d, p0, seed, counter, out):
        self._check(self.lib.so_swarm_rand(self.handle, int(P), int(d), int(p0), int(seed), int(counter), _ptr(out), self._stream()),
                    "so_swarm_rand")
        self.launches += 1

    def swarm_step_dev(self, pos, vel, best_pos, global_best, state, seed, p0, velocity_scale, bounds):
        P, d = pos.shape
        vs = _np_f64(velocity_scale)
        bd = None if bounds is None else _np_f64(bounds)
        rc = self.lib.so_swarm_step_dev(self.handle, P, d, int(p0), _ptr(pos), _ptr(vel), _ptr(best_pos), _ptr(global_best),
                                        _ptr(state), int(seed), _hptr(vs), _hptr(bd), self._stream())
        self._check(rc, "so_swarm_step_dev")
        self.launches += 1

    def swarm_update_best_x(self, pos, values, safe, best_pos, best_values, best_idx, p0, global_best, global_rec, state=None):
        P, d = pos.shape
        rc = self.lib.so_swarm_update_best_x(self.handle, P, d, _ptr(pos), _ptr(values), _ptr(safe), _ptr(best_pos),
                                             _ptr(best_values), _ptr(best_idx), int(p0), _ptr(global_best), _ptr(global_rec),
                                             _ptr(state), self._stream())
        self._check(rc, "so_swarm_update_best_x")
        self.launches += 1

    def swarm_combine_best(self, recs, d, global_best, global_rec=None):
        rc = self.lib.so_swarm_combine_best(self.handle, _ptr(recs), int(recs.shape[0]), int(d), _ptr(global_best),
                                            _ptr(global_rec), self._stream())
        self._check(rc, "so_swarm_combine_best")
        self.launches += 1

    # ------------------------------------------------------------------ f1: swarm safe-set maintenance
    def safeset_filter(self, gp, cand, ref, scale2, thresh, keep):
        n, m = cand.shape[0], ref.shape[0]
        rc = self.lib.so_safeset_filter(self.handle, gp, _ptr(cand), n, _ptr(ref) if m else C.c_void_p(0), m, float(scale2),
                                        float(thresh), _ptr(keep), self._stream())
        self._check(rc, "so_safeset_filter")
        self.launches += (1 + (1 if m else 0)) if n else 0

    def safeset_insert(self, gp, cand, keep, scale2, thresh, accept, accepted_pos, n_accept):
        n = cand.shape[0]
        rc = self.lib.so_safeset_insert(self.handle, gp, _ptr(cand), n, _ptr(keep), float(scale2), float(thresh),
                                        _ptr(accept), _ptr(accepted_pos), _ptr(n_accept), self._stream())
        self._check(rc, "so_safeset_insert")
        blocks = (n + 1023) // 1024
        self.launches += max(0, 2 * blocks - 1)

    # ------------------------------------------------------------------ records
    def read_record(self, rec, dtype):
        return rec.cpu().numpy().view(dtype)[0]
