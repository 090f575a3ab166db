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

import gc
import os

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
    a sharded run follows the single-GPU trajectory exactly) or 'device' (counter-based Philox
    randoms drawn inside the update kernel, keyed by ``seed``, the iteration number and the GLOBAL
    particle index -- no host traffic, and again independent of the sharding).

    With ``torch.distributed`` initialised the P particles are split into contiguous blocks
    (``distributed.shard_bounds``).  One PSO iteration is: K6 update (randoms drawn in the kernel, or
    copied from the host) -> posterior + fitness -> one kernel that updates the personal bests, writes
    this rank's {value, global index, position} record into every rank's exchange buffer over NVLink,
    waits for the ``world`` records of the iteration and combines the global best
    (``so_swarm_update_best_x``).  Nothing in the iteration waits for the host; with ``rng='device'``
    nothing in its launches changes between iterations either (inertia and iteration number live in
    device memory), so ``run_swarm`` captures ONE iteration in a CUDA graph and replays it
    (``SAFEOPT_B200_SWARM_GRAPH=0`` launches kernel by kernel).  Where the peer mapping is unavailable
    the records go through one all-gather per iteration instead (and no graph).
    """

    def __init__(self, engine, velocity, fitness_device, bounds=None, rng="host", seed=0, comm=None, peer=None):
        from . import _lib
        from .distributed import Comm
        self.engine = engine
        self.velocity_scale = np.asarray(velocity, dtype=float)
        self.ndim = len(self.velocity_scale)
        self.fitness_device = fitness_device
        self.bounds = None if bounds is None else np.ascontiguousarray(np.asarray(bounds, dtype=float))
        self.initial_inertia, self.final_inertia = 1.0, 0.1
        if rng not in ("host", "device"):
            raise ValueError("rng must be 'host' or 'device'")
        self.rng = rng
        self.seed = int(seed)
        self.comm = comm if comm is not None else Comm(engine.device)
        if peer is None:
            peer = bool(hasattr(engine, "connect_exchange") and engine.connect_exchange(self.comm))
        # records exchanged by the kernel itself (peer-mapped memory); otherwise one all-gather per iteration
        self.in_kernel_exchange = hasattr(engine, "swarm_update_best_x") and (peer or not self.comm.active)
        self.use_graph = (self.in_kernel_exchange and rng == "device" and engine.device.type == "cuda"
                          and os.environ.get("SAFEOPT_B200_SWARM_GRAPH", "1") != "0")
        self.fitness_key = None             # callable -> hashable: what the fitness callback bakes into its launches
        self._graph = None
        self._graph_key = None
        self._graph_launches = 0
        self._side = None                   # capture stream
        self._draws = 0                     # random blocks drawn so far outside the update kernel (rng='device')
        self.swarm_size = 0                 # particles of the whole swarm
        self.p0 = self.p1 = 0               # this rank's block
        self.positions = self.velocities = self.best_positions = self.best_values = None
        self._best_idx = engine.zeros((1,), "i64")
        self._rec = engine.zeros((_lib.SWARM_REC_DOUBLES,))
        self._recs = engine.zeros((self.comm.world, _lib.SWARM_REC_DOUBLES))
        self._grec = engine.zeros((4,))
        self._state = engine.zeros((4,))    # inertia, inertia step, iteration number (device rng / graph mode)
        self.global_best_d = engine.zeros((self.ndim,))

    # ---- host views
    @property
    def global_best(self):
        """Best position found so far (host copy)."""
        return self.global_best_d.cpu().numpy()

    @property
    def global_best_value(self):
        g = self._grec.cpu().numpy()
        if g[2] != 0.0:
            raise RuntimeError("swarm best-record exchange timed out: a rank never published its record")
        return float(g[0])

    @property
    def max_velocity(self):
        return 10 * self.velocity_scale

    def _rand_rows(self, blocks):
        """``blocks`` stacked (P_total, d) uniform blocks as the reference draws them; returns this rank's
        rows of each block, stacked, on the device."""
        n_local = self.p1 - self.p0
        if self.rng == "device":
            out = self.engine.empty((blocks * n_local, self.ndim))
            for b in range(blocks):
                # counters count down from 2^63 so that they never meet the iteration numbers the update kernel uses
                self.engine.swarm_rand(n_local, self.ndim, self.p0, self.seed, (1 << 63) - 1 - self._draws,
                                       out[b * n_local:(b + 1) * n_local])
                self._draws += 1
            return out
        from .distributed import shard_stacked_blocks
        full = np.random.rand(blocks * self.swarm_size, self.ndim)
        return self.engine.to_device(shard_stacked_blocks(full, blocks, self.swarm_size, self.p0, self.p1))

    def _update_best(self, values, safe, state=None):
        """Personal bests, this rank's record, exchange, global best."""
        eng = self.engine
        if self.in_kernel_exchange:
            eng.swarm_update_best_x(self.positions, values, safe, self.best_positions, self.best_values, self._best_idx,
                                    self.p0, self.global_best_d, self._grec, state)
            return
        eng.swarm_update_best(self.positions, values, safe, self.best_positions, self.best_values, self._best_idx,
                              p0=self.p0, rec=self._rec)
        self.comm.all_gather_into(self._recs, self._rec)
        eng.swarm_combine_best(self._recs, self.ndim, self.global_best_d, self._grec)

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
        if self.positions is not None and tuple(self.positions.shape) == tuple(local.shape):
            # same block size as the last swarm: keep the buffers (their addresses are baked into the captured iteration)
            self.positions.copy_(local if t.is_tensor(local) else t.from_numpy(np.ascontiguousarray(local)))
            self.velocities.copy_(self._rand_rows(1) * eng.to_device(self.velocity_scale))
        else:
            self.positions = local.clone() if t.is_tensor(local) else eng.to_device(np.ascontiguousarray(local))
            self.velocities = self._rand_rows(1) * eng.to_device(self.velocity_scale)
            self.best_positions = eng.empty(tuple(self.positions.shape))
            self.best_values = eng.empty((self.p1 - self.p0,))
            self._graph = None
        values, _ = self.fitness_device(self.positions)
        self.best_positions.copy_(self.positions)
        self.best_values.copy_(values)
        # first-index argmax on the device, via the same kernel that tracks bests (safety is ignored at init, swarm.py:78-84)
        never = eng.zeros((self.p1 - self.p0,), "u8")
        self._update_best(values, never)

    def _iteration_dev(self):
        """One PSO iteration whose launches are identical from one iteration to the next (device rng)."""
        eng = self.engine
        eng.swarm_step_dev(self.positions, self.velocities, self.best_positions, self.global_best_d, self._state, self.seed,
                           self.p0, self.velocity_scale, self.bounds)
        values, safe = self.fitness_device(self.positions)
        self._update_best(values, safe, self._state)

    def run_swarm(self, max_iter):
        eng = self.engine
        inertia = self.initial_inertia
        step = (self.final_inertia - self.initial_inertia) / max_iter
        if self.rng == "device" and self.in_kernel_exchange:
            t = eng.torch
            it0 = float(self._state[2].item())       # the iteration number keeps counting over runs: fresh randoms every run
            self._state.copy_(t.tensor([inertia, step, it0, 0.0], dtype=t.float64))
            if not self.use_graph:
                for _ in range(max_iter):
                    self._iteration_dev()
                return
            # everything a captured launch bakes in: buffer addresses here, and whatever the fitness callback passes by value
            # (beta, thresholds, the fit it evaluates) -- the owner describes that with `fitness_key`
            key = (self.positions.data_ptr(), self.velocities.data_ptr(), self.best_positions.data_ptr(), self.best_values.data_ptr(),
                   self.fitness_key() if self.fitness_key is not None else None)
            done = 0
            if self._graph is None or self._graph_key != key:
                # one iteration outside the graph first (kernel attributes are set and scratch is allocated on first use), then
                # capture; the fitness callback must not allocate or synchronise (SafeOptSwarm._swarm_fitness does not)
                self._iteration_dev()
                done = 1
                # capture_begin / capture_end directly: the torch.cuda.graph() context also runs gc.collect() and empty_cache(),
                # tens of milliseconds per capture, and a SafeOptSwarm.optimize() re-captures three times (new greedy bound)
                g = t.cuda.CUDAGraph()
                cur = t.cuda.current_stream(eng.device)
                if self._side is None:
                    self._side = t.cuda.Stream(device=eng.device)
                side = self._side
                side.wait_stream(cur)
                before = eng.launches
                # No cyclic garbage collection while the stream is capturing: a collection may finalise an engine of an earlier
                # optimiser (SafeOptSwarm <-> DeviceSwarm is a reference cycle, so such objects die only in the collector), its
                # so_destroy() calls cudaFree / cudaFreeHost, and those are forbidden during a capture -- the capture is
                # invalidated and capture_end raises (seen as an intermittent torch.AcceleratorError that depended on what ran
                # before).  torch.cuda.graph() avoids the same thing by collecting up front, at tens of milliseconds per capture.
                gc_was_enabled = gc.isenabled() and os.environ.get("SAFEOPT_B200_GC_IN_CAPTURE", "0") != "1"
                if gc_was_enabled:
                    gc.disable()
                try:
                    with t.cuda.stream(side):
                        g.capture_begin()
                        try:
                            self._iteration_dev()
                        finally:
                            g.capture_end()
                finally:
                    if gc_was_enabled:
                        gc.enable()
                cur.wait_stream(side)
                # capturing does not execute: the captured iteration has not run yet
                self._graph_launches = eng.launches - before
                eng.launches = before
                self._graph, self._graph_key = g, key
            for _ in range(max_iter - done):
                self._graph.replay()
            eng.launches += self._graph_launches * (max_iter - done)
            return
        for _ in range(max_iter):
            r = self._rand_rows(2)
            eng.swarm_step(self.positions, self.velocities, self.best_positions, self.global_best_d, r, inertia,
                           self.velocity_scale, self.bounds)
            inertia += step
            values, safe = self.fitness_device(self.positions)
            self._update_best(values, safe)
