"""Row-block sharding of the candidate set and the tiny cross-rank reductions.

One process per GPU (``torchrun``); the training-side fit is recomputed redundantly on every
rank (deterministic kernels => bit-identical), the M candidate rows are split into contiguous
blocks, and only 64-byte records cross NVLink: ``(any, max, (value,row))`` all-gathers.  NCCL
has no MAXLOC, hence all-gather + a local, deterministic combine with lowest-global-row
tie-breaks (NumPy's first-occurrence argmax).

Works without ``torch.distributed`` initialised (world size 1) and with the ``gloo`` backend on
CPU tensors (used by the CPU test-suite for the host-side logic).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank``: blocks of ceil(M/R) rows, the last ones possibly short/empty."""
    per = -(-int(n_rows) // int(world))
    lo = min(rank * per, n_rows)
    hi = min(lo + per, n_rows)
    return lo, hi


class Comm:
    """Thin wrapper over torch.distributed that degrades to a single rank."""

    def __init__(self, device=None):
        self.rank, self.world = 0, 1
        self._dist = None
        self._device = device
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank = dist.get_rank()
                self.world = dist.get_world_size()
        except ImportError:  # pragma: no cover
            pass

    @property
    def active(self) -> bool:
        return self.world > 1

    def _tensor_device(self):
        if self._dist is None:
            return None
        backend = self._dist.get_backend()
        if backend == "nccl":
            import torch
            return self._device if self._device is not None else torch.device("cuda", torch.cuda.current_device())
        return "cpu"

    def all_gather(self, arr: np.ndarray) -> np.ndarray:
        """Gather equally-shaped float64/int64/uint8 arrays; returns shape (world, *arr.shape)."""
        arr = np.ascontiguousarray(arr)
        if not self.active:
            return arr[None, ...]
        import torch
        dev = self._tensor_device()
        t = torch.from_numpy(arr).to(dev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self._dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out], axis=0)

    def any_flags(self, flags: np.ndarray) -> np.ndarray:
        """Elementwise OR of uint8 flag arrays across ranks."""
        if not self.active:
            return flags
        return self.all_gather(flags.astype(np.uint8)).max(axis=0)

    def barrier(self):
        if self.active:
            self._dist.barrier()


def combine_max_first(values: np.ndarray, rows: np.ndarray) -> Tuple[float, int]:
    """Reduce per-rank (value, global_row) pairs: largest value, ties -> smallest row; rows < 0 are empty."""
    best_v, best_r = -np.inf, -1
    for v, r in zip(np.asarray(values).ravel(), np.asarray(rows).ravel()):
        r = int(r)
        if r < 0:
            continue
        if best_r < 0 or v > best_v or (v == best_v and r < best_r):
            best_v, best_r = float(v), r
    return best_v, best_r
