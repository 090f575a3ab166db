"""A NumPy stand-in for safeopt_b200.engine.DeviceEngine, built on the oracle.  TEST INFRASTRUCTURE ONLY.

It lets the CPU suite drive the product's *host orchestration* (safeopt_b200/gp_opt.py: fit bookkeeping, the chained set
passes and their record parsing, candidate ordering, the batched expander search, query-point selection) against the golden
fixtures without a GPU.  Every method mirrors the contract of the C entry point of the same name in
include/safeopt_b200.h; buffers are torch CPU tensors so that the product code's tensor plumbing runs unchanged.
The product never imports this module.
"""
import numpy as np
import torch

from oracle import gpy_lite, safeopt_port as port
from safeopt_b200 import _lib
from safeopt_b200.engine import MAX_REC_DTYPE, SAFE_REC_DTYPE
from safeopt_b200.utilities import grid_rows_from_index

_DT = {"f64": torch.float64, "u8": torch.uint8, "i64": torch.int64, "f32": torch.float32}
_KERN = {0: gpy_lite.RBF, 1: gpy_lite.Matern32, 2: gpy_lite.Matern52}


class FakeEngine:
    def __init__(self, device=None, max_gps=8):
        self.torch = torch
        self.device = torch.device("cpu")
        self.max_gps = max_gps
        self.launches = 0
        self.gps = {}
        self.axes = None
        self.calls = []

    # ---- buffers
    def empty(self, shape, dtype="f64"):
        return torch.zeros(shape, dtype=_DT[dtype])

    zeros = empty

    def to_device(self, arr, pinned=False):
        return torch.from_numpy(np.ascontiguousarray(arr)).clone()

    def synchronize(self):
        pass

    def close(self):
        pass

    # ---- K1
    def fit(self, gp, X, Y, kind, lengthscale, variance, noise_var):
        X = np.asarray(X, dtype=float)
        ls = np.broadcast_to(np.asarray(lengthscale, dtype=float).reshape(-1), (X.shape[1],)).copy()
        kern = _KERN[int(kind)](X.shape[1], variance=float(variance), lengthscale=ls, ARD=True)
        self.gps[gp] = gpy_lite.GPRegression(X, np.asarray(Y, dtype=float).reshape(-1, 1), kernel=kern, noise_var=float(noise_var))
        self.calls.append(("fit", gp, X.shape[0]))

    def fit_append(self, gp, x_new, y_new):
        g = self.gps[gp]
        g.set_XY(np.vstack([g.X, np.asarray(x_new, dtype=float).reshape(1, -1)]), np.vstack([g.Y, [[float(y_new)]]]))
        self.calls.append(("append", gp))
        return True

    def fit_remove_last(self, gp):
        g = self.gps[gp]
        g.set_XY(g.X[:-1], g.Y[:-1])
        self.calls.append(("remove", gp))

    # ---- grid
    def define_grid(self, axes):
        self.axes = [np.asarray(a, dtype=float) for a in axes]
        self._grid_axes = self.axes

    def prepare_grid(self, gp, row0=None, n_rows=None):
        self.calls.append(("prepare", gp, row0, n_rows))

    def grid_rows(self, row0, M):
        return torch.from_numpy(grid_rows_from_index(self.axes, np.arange(row0, row0 + M)))

    # ---- K2
    def _rows(self, Xstar, row0, M):
        return grid_rows_from_index(self.axes, np.arange(row0, row0 + M)) if Xstar is None else Xstar.numpy()[:M]

    def _posterior(self, gps, X, beta, fmins, means, variances, Q, q_cols, S, safe_mode):
        safe = np.ones(X.shape[0], dtype=bool)
        for k, gp in enumerate(gps):
            m, v = self.gps[gp].predict_noiseless(X)
            m, v = m[:, 0], v[:, 0]
            sd = np.sqrt(v)
            lo, up = m - beta * sd, m + beta * sd
            if means is not None and means[k] is not None:
                means[k].numpy()[:] = m
            if variances is not None and variances[k] is not None:
                variances[k].numpy()[:] = v
            if Q is not None:
                Q.numpy()[:, q_cols[k]] = lo
                Q.numpy()[:, q_cols[k] + 1] = up
            safe &= lo > fmins[k]
        if S is not None and safe_mode != _lib.SAFE_NONE:
            s = S.numpy()
            s[:] = safe if safe_mode == _lib.SAFE_WRITE else (s.astype(bool) & safe)
        self.launches += 1

    def posterior_rows(self, gp, Xstar, beta, fmin, mean=None, var=None, Q=None, q_col=0, S=None, safe_mode=_lib.SAFE_NONE):
        self._posterior([gp], Xstar.numpy(), beta, [fmin], [mean], [var], Q, [q_col], S, safe_mode)

    def posterior_grid(self, gp, row0, M, beta, fmin, mean=None, var=None, Q=None, q_col=0, S=None, safe_mode=_lib.SAFE_NONE):
        self._posterior([gp], self._rows(None, row0, M), beta, [fmin], [mean], [var], Q, [q_col], S, safe_mode)

    def posterior_multi(self, gps, Xstar, row0, M, beta, fmins, means=None, variances=None, Q=None, q_cols=None, S=None,
                        safe_mode=_lib.SAFE_NONE):
        self._posterior(list(gps), self._rows(Xstar, row0, M), beta, list(fmins), means, variances, Q,
                        list(q_cols) if q_cols is not None else [0] * len(gps), S, safe_mode)
        self.calls.append(("multi", tuple(gps)))
        return True

    # ---- K3
    @staticmethod
    def _first_max(values, rows):
        if rows.size == 0:
            return -np.inf, -1
        k = int(np.argmax(values))
        return float(values[k]), int(rows[k])

    def reduce_safe(self, Q, n_gps, row0, S, rec):
        q, s = Q.numpy(), S.numpy().astype(bool)
        rows = np.flatnonzero(s)
        out = np.zeros(1, dtype=SAFE_REC_DTYPE)
        out["n_safe"] = rows.size
        out["max_l0"], out["argmax_l0"] = self._first_max(q[rows, 0], rows + row0)
        out["max_u0"], out["argmax_u0"] = self._first_max(q[rows, 1], rows + row0)
        rec.numpy().reshape(-1)[:64] = out.view(np.uint8)
        self.launches += 1

    def maximizers_chain(self, Q, n_gps, row0, S, safe_recs, n_recs, scaling, Mmask, rec):
        recs = np.ascontiguousarray(safe_recs.numpy()).reshape(-1)[:64 * n_recs].view(SAFE_REC_DTYPE)
        max_l0 = float(np.max(recs["max_l0"]))
        q, s = Q.numpy(), S.numpy().astype(bool)
        m = s & (q[:, 1] >= max_l0)
        Mmask.numpy()[:] = m
        rows = np.flatnonzero(m)
        out = np.zeros(1, dtype=MAX_REC_DTYPE)
        out["n_max"] = rows.size
        width = q[rows, 1::2] - q[rows, ::2]
        out["max_width0"] = width[:, 0].max() if rows.size else -np.inf
        out["best_value"], out["best_row"] = self._first_max(np.max(width / np.asarray(scaling), axis=1) if rows.size else np.zeros(0),
                                                             rows + row0)
        rec.numpy().reshape(-1)[:64] = out.view(np.uint8)
        self.launches += 1

    def candidates_chain(self, Q, n_gps, row0, S, Mmask, max_recs, n_recs, scaling, thr, cand_mask, cand_key, cand_row, n_cand):
        recs = np.ascontiguousarray(max_recs.numpy()).reshape(-1)[:64 * n_recs].view(MAX_REC_DTYPE)
        max_var = float(np.max(recs["max_width0"])) / float(np.asarray(scaling)[0])
        q = Q.numpy()
        width = q[:, 1::2] - q[:, ::2]
        c = S.numpy().astype(bool) & ~Mmask.numpy().astype(bool)
        c &= np.max(width / np.asarray(scaling), axis=1) > max_var
        c &= np.any(width > np.asarray(thr), axis=1)
        rows = np.flatnonzero(c)[::-1]                      # the device appends in no particular order
        n_cand.numpy()[0] = rows.size
        cand_key.numpy()[:rows.size] = np.max(width[rows], axis=1)
        cand_row.numpy()[:rows.size] = rows + row0
        self.launches += 1

    # ---- K4
    def expander_check(self, gp, Xstar, row0, M, S, mean, var, xc, mean_c, var_c, u_c, beta, fmin, flags):
        g = self.gps[gp]
        X = self._rows(Xstar, row0, M)
        unsafe = ~S.numpy().astype(bool)
        Xu, mu, vu = X[unsafe], mean.numpy()[unsafe], var.numpy()[unsafe]
        W = g.woodbury_inv
        Kxu = g.kern.K(g.X, Xu)
        noise = float(g.likelihood.variance) + gpy_lite.JITTER
        for b in range(xc.shape[0]):
            kc = g.kern.K(g.X, xc.numpy()[[b]])[:, 0]
            c = g.kern.K(Xu, xc.numpy()[[b]])[:, 0] - Kxu.T @ (W @ kc)
            s = float(var_c.numpy()[b]) + noise
            m2 = mu + c * (float(u_c.numpy()[b]) - float(mean_c.numpy()[b])) / s
            v2 = np.maximum(vu - c * c / s, gpy_lite.VAR_FLOOR)
            if Xu.shape[0] and np.any(m2 - beta * np.sqrt(v2) >= fmin):
                flags.numpy()[b] |= 1
        self.launches += 2

    def expander_lipschitz(self, Xstar, d, row0, M, S, xc, u_c, lipschitz, fmin, flags):
        X = self._rows(Xstar, row0, M)
        Xu = X[~S.numpy().astype(bool)]
        for b in range(xc.shape[0]):
            dist = np.sqrt(np.sum((Xu - xc.numpy()[b]) ** 2, axis=1))
            if Xu.shape[0] and np.any(float(u_c.numpy()[b]) - lipschitz * dist >= fmin):
                flags.numpy()[b] |= 1
        self.launches += 1

    # ---- swarm fitness epilogue and safe-set maintenance (host-backend SafeOptSwarm)
    def swarm_fitness(self, kind, n_gps, P, mean, var, beta, fmin, scaling, best_lower_bound, values, safe):
        from scipy.special import expit
        from scipy.stats import norm
        name = {v: k for k, v in _lib.SWARM_KINDS.items()}[kind]
        m, v = mean.numpy()[:, :P], var.numpy()[:, :P]
        sd = np.sqrt(v[0])
        lower, upper = m[0] - beta * sd, m[0] + beta * sd
        if name == "greedy":
            values.numpy()[:] = lower
            safe.numpy()[:] = 1
            return
        vals = sd / scaling[0]
        interest = {"safe_set": None, "expanders": n_gps * np.ones(P),
                    "maximizers": expit(10 * (upper - best_lower_bound) / scaling[0])}[name]
        ok, pen = np.ones(P, dtype=bool), np.zeros(P)
        for i in range(n_gps):
            if i > 0:
                sd = np.sqrt(v[i])
                lower = m[i] - beta * sd
                vals = np.maximum(vals, sd / scaling[i])
            if fmin[i] == -np.inf:
                continue
            slack = lower - fmin[i]
            ok &= slack >= 0
            if name == "safe_set":
                continue
            slack = slack / scaling[i]
            pen += port.penalty(slack)
            if name == "expanders":
                interest = interest * norm.pdf(slack, scale=0.2)
        values.numpy()[:] = lower if name == "safe_set" else (vals + pen) * interest
        safe.numpy()[:] = ok
        self.launches += 1

    def safeset_filter(self, gp, cand, ref, scale2, thresh, keep):
        g = self.gps[gp]
        c = cand.numpy()
        keep.numpy()[:] = np.all(g.kern.K(c, ref.numpy()) / scale2 <= thresh, axis=1) if ref.shape[0] else 1

    def safeset_insert(self, gp, cand, keep, scale2, thresh, accept, accepted_pos, n_accept):
        g = self.gps[gp]
        c, k = cand.numpy(), keep.numpy().astype(bool)
        acc = np.zeros(c.shape[0], dtype=bool)
        for j in np.flatnonzero(k):
            if not acc.any() or np.all(g.kern.K(c[[j]], c[acc])[0] / scale2 <= thresh):
                acc[j] = True
        accept.numpy()[:] = acc
        accepted_pos.numpy()[:acc.sum()] = c[acc]
        n_accept.numpy()[0] = acc.sum()

    # ---- device swarm (K6): update rule of safeopt/swarm.py:98-146 on the stand-in tensors
    def swarm_step(self, pos, vel, best_pos, global_best, r, inertia, velocity_scale, bounds):
        P = pos.shape[0]
        x, v, bp, gb, rr = pos.numpy(), vel.numpy(), best_pos.numpy(), global_best.numpy(), r.numpy()
        vs = np.asarray(velocity_scale, dtype=float)
        v *= inertia
        v += (rr[:P] * (bp - x) + rr[P:] * (gb - x)) / vs
        np.clip(v, -10 * vs, 10 * vs, out=v)
        x += v
        if bounds is not None:
            b = np.asarray(bounds, dtype=float)
            np.clip(x, b[:, 0], b[:, 1], out=x)
        self.launches += 1

    def swarm_update_best(self, pos, values, safe, best_pos, best_values, best_idx, p0=0, rec=None):
        x, val, ok = pos.numpy(), values.numpy(), safe.numpy().astype(bool)
        bp, bv = best_pos.numpy(), best_values.numpy()
        better = (val > bv) & ok
        bv[better] = val[better]
        bp[better] = x[better]
        k = int(np.argmax(bv))
        best_idx.numpy()[0] = k
        if rec is not None:
            out = rec.numpy()
            out[0], out[1] = bv[k], float(p0 + k)
            out[2:2 + x.shape[1]] = bp[k]
        self.launches += 1

    def swarm_combine_best(self, recs, d, global_best, global_rec=None):
        r = recs.numpy()
        best = None
        for row in r:
            if row[1] < 0:
                continue
            if best is None or row[0] > best[0] or (row[0] == best[0] and row[1] < best[1]):
                best = row
        global_best.numpy()[:] = best[2:2 + d]
        if global_rec is not None:
            global_rec.numpy()[:2] = best[:2]
        self.launches += 1
