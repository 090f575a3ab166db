"""Constrained particle-swarm optimisation (reference: /root/reference/safeopt/swarm.py:17-146).

Two drivers with the same update rule (c1 = c2 = 1, inertia 1.0 -> 0.1 linearly, velocity clipped
to +-10 * velocity_scale, positions clipped to the bounds, personal bests updated only where the
new value is better AND safe):

* :class:`SwarmOptimization` -- the reference's class surface: host NumPy state, randoms from the
  global ``np.random`` stream (so trajectories are reproducible against the reference for the
  same seed), ``fitness(positions) -> (values, safe)`` callback.  With SafeOptSwarm the callback
  evaluates the GP posterior of all particles on the GPU.
* :class:`DeviceSwarm` -- for large swarms (BASELINE config 5: 1e5 particles): positions,
  velocities and bests stay in HBM, the update and best-tracking are the K6 kernels, only the
  uniform randoms come from the host (or from torch's device generator when ``rng='device'``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["SwarmOptimization", "DeviceSwarm"]


class SwarmOptimization(object):
    """Constrained swarm optimisation.

    Parameters
    ----------
    swarm_size : int
    velocity : ndarray -- base velocity per dimension.
    fitness : callable(positions) -> (values, safe_mask)
    bounds : list of (min, max), optional.
    """

    def __init__(self, swarm_size, velocity, fitness, bounds=None):
        super(SwarmOptimization, self).__init__()
        self.c1 = self.c2 = 1
        self.fitness = fitness
        self.bounds = None if bounds is None else np.asarray(bounds)
        self.initial_inertia = 1.0
        self.final_inertia = 0.1
        self.velocity_scale = velocity
        self.ndim = len(velocity)
        self.swarm_size = swarm_size
        self.positions = np.empty((swarm_size, self.ndim), dtype=float)
        self.velocities = np.empty_like(self.positions)
        self.best_positions = np.empty_like(self.positions)
        self.best_values = np.empty(swarm_size, dtype=float)
        self.global_best = None

    @property
    def max_velocity(self):
        """Largest allowed particle speed per dimension."""
        return 10 * self.velocity_scale

    def init_swarm(self, positions):
        """Place the particles, draw initial velocities, score once (swarm.py:66-84)."""
        self.positions = positions
        self.velocities = np.random.rand(*self.velocities.shape) * self.velocity_scale
        values, _ = self.fitness(self.positions)
        self.best_positions[:] = self.positions
        self.best_values = values
        self.global_best = self.best_positions[np.argmax(values), :]

    def run_swarm(self, max_iter):
        """Iterate the swarm ``max_iter`` times (swarm.py:86-146)."""
        inertia = self.initial_inertia
        step = (self.final_inertia - self.initial_inertia) / max_iter
        for _ in range(max_iter):
            to_global = self.global_best - self.positions
            to_self = self.best_positions - self.positions
            r = np.random.rand(2 * self.swarm_size, self.ndim)
            r_self, r_global = r[:self.swarm_size], r[self.swarm_size:]
            self.velocities *= inertia
            self.velocities += (self.c1 * r_self * to_self + self.c2 * r_global * to_global) / self.velocity_scale
            inertia += step
            np.clip(self.velocities, -self.max_velocity, self.max_velocity, out=self.velocities)
            self.positions += self.velocities
            if self.bounds is not None:
                np.clip(self.positions, self.bounds[:, 0], self.bounds[:, 1], out=self.positions)
            values, safe = self.fitness(self.positions)
            better = (values > self.best_values) & safe
            self.best_values[better] = values[better]
            self.best_positions[better] = self.positions[better]
            self.global_best = self.best_positions[np.argmax(self.best_values), :]


class DeviceSwarm(object):
    """Device-resident swarm for large particle counts, optionally sharded over ranks.

    ``fitness_device(positions_tensor) -> (values_tensor, safe_u8_tensor)`` must keep everything on
    the GPU (``SafeOptSwarm._fitness_device`` does).  ``rng`` is 'host' (NumPy global stream, the
    reference's source of randomness; every rank draws the full-swarm block and keeps its rows, so
    a sharded run follows the single-GPU trajectory exactly) or 'device' (torch generator seeded
    ``seed + rank``, no host traffic).

    With ``torch.distributed`` initialised the P particles are split into contiguous blocks
    (``distributed.shard_bounds``); one PSO iteration is: randoms -> K6 update -> posterior + fitness
    -> personal bests + this rank's {value, global index, position} record (one kernel) ->
    all-gather of the 144-byte records -> combine kernel -> ``global_best`` in device memory.
    Nothing in the iteration waits for the host (one record all-gather per iteration).
    """

    def __init__(self, engine, velocity, fitness_device, bounds=None, rng="host", seed=0, comm=None):
        from . import _lib
        from .distributed import Comm
        self.engine = engine
        self.velocity_scale = np.asarray(velocity, dtype=float)
        self.ndim = len(self.velocity_scale)
        self.fitness_device = fitness_device
        self.bounds = None if bounds is None else np.ascontiguousarray(np.asarray(bounds, dtype=float))
        self.initial_inertia, self.final_inertia = 1.0, 0.1
        self.rng = rng
        self.comm = comm if comm is not None else Comm(engine.device)
        self._gen = None
        if rng == "device":
            self._gen = engine.torch.Generator(device=engine.device)
            self._gen.manual_seed(seed + self.comm.rank)
        self.swarm_size = 0                 # particles of the whole swarm
        self.p0 = self.p1 = 0               # this rank's block
        self.positions = self.velocities = self.best_positions = self.best_values = None
        self._best_idx = engine.zeros((1,), "i64")
        self._rec = engine.zeros((_lib.SWARM_REC_DOUBLES,))
        self._recs = engine.zeros((self.comm.world, _lib.SWARM_REC_DOUBLES))
        self._grec = engine.zeros((2,))
        self.global_best_d = engine.zeros((self.ndim,))

    # ---- host views
    @property
    def global_best(self):
        """Best position found so far (host copy)."""
        return self.global_best_d.cpu().numpy()

    @property
    def global_best_value(self):
        return float(self._grec[0].item())

    @property
    def max_velocity(self):
        return 10 * self.velocity_scale

    def _rand_rows(self, blocks):
        """``blocks`` stacked (P_total, d) uniform blocks as the reference draws them; returns this rank's
        rows of each block, stacked, on the device."""
        t = self.engine.torch
        n_local = self.p1 - self.p0
        if self.rng == "device":
            return t.rand((blocks * n_local, self.ndim), dtype=t.float64, device=self.engine.device, generator=self._gen)
        from .distributed import shard_stacked_blocks
        full = np.random.rand(blocks * self.swarm_size, self.ndim)
        return self.engine.to_device(shard_stacked_blocks(full, blocks, self.swarm_size, self.p0, self.p1))

    def _exchange_best(self):
        """Per-rank best records -> ``global_best_d`` (all ranks end up with the same bytes)."""
        self.comm.all_gather_into(self._recs, self._rec)
        self.engine.swarm_combine_best(self._recs, self.ndim, self.global_best_d, self._grec)

    def init_swarm(self, positions):
        """``positions``: the whole swarm (P, d), identical on every rank (host array or device tensor)."""
        from .distributed import shard_bounds
        eng = self.engine
        t = eng.torch
        self.swarm_size = int(positions.shape[0])
        self.p0, self.p1 = shard_bounds(self.swarm_size, self.comm.world, self.comm.rank)
        if self.p1 <= self.p0:
            raise ValueError("swarm of %d particles cannot be split over %d ranks" % (self.swarm_size, self.comm.world))
        local = positions[self.p0:self.p1]
        self.positions = local.clone() if t.is_tensor(local) else eng.to_device(np.ascontiguousarray(local))
        self.velocities = self._rand_rows(1) * eng.to_device(self.velocity_scale)
        values, _ = self.fitness_device(self.positions)
        self.best_positions = self.positions.clone()
        self.best_values = values.clone()
        # first-index argmax on the device, via the same kernel that tracks bests (safety is ignored at init, swarm.py:78-84)
        never = eng.zeros((self.p1 - self.p0,), "u8")
        eng.swarm_update_best(self.positions, values, never, self.best_positions, self.best_values, self._best_idx,
                              p0=self.p0, rec=self._rec)
        self._exchange_best()

    def run_swarm(self, max_iter):
        eng = self.engine
        inertia = self.initial_inertia
        step = (self.final_inertia - self.initial_inertia) / max_iter
        for _ in range(max_iter):
            r = self._rand_rows(2)
            eng.swarm_step(self.positions, self.velocities, self.best_positions, self.global_best_d, r, inertia,
                           self.velocity_scale, self.bounds)
            inertia += step
            values, safe = self.fitness_device(self.positions)
            eng.swarm_update_best(self.positions, values, safe, self.best_positions, self.best_values, self._best_idx,
                                  p0=self.p0, rec=self._rec)
            self._exchange_best()
