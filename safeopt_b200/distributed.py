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


def torch_empty_like(t):
    import torch
    return torch.empty_like(t)


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank``: blocks of ceil(M/R) rows, the last ones possibly short/empty."""
    per = -(-int(n_rows) // int(world))
    lo = min(rank * per, n_rows)
    hi = min(lo + per, n_rows)
    return lo, hi


class Comm:
    """Thin wrapper over torch.distributed that degrades to a single rank."""

    def __init__(self, device=None, enabled=True):
        self.rank, self.world = 0, 1
        self._dist = None
        self._device = device
        self._flat_gather = False
        if not enabled:                 # a deliberately rank-local object inside a multi-rank job
            return
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank = dist.get_rank()
                self.world = dist.get_world_size()
                self._flat_gather = hasattr(dist, "all_gather_into_tensor")
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
        return self.all_gather_tensor(t).reshape((self.world,) + arr.shape)

    def all_gather_tensor(self, t):
        """Gather a (device) tensor from every rank with ONE collective and ONE copy to the host;
        returns a NumPy array of shape (world, *t.shape).  Used for the 64-byte records the set
        kernels leave in device memory, so a record costs gather + D2H instead of D2H + H2D + gather + D2H."""
        if not self.active:
            return t.cpu().numpy()[None, ...]
        import torch
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.all_gather_into(out, t)
        return out.cpu().numpy()

    def all_gather_into(self, out, t):
        """Device-side all-gather with no host involvement: ``out[r] = t`` of rank r (``out``: (world, *t.shape))."""
        if not self.active:
            out[0].copy_(t)
            return out
        # the variant is chosen ONCE by capability, never by catching an error of a live collective (a retry with a
        # different collective after a genuine failure would desynchronise the ranks)
        if self._flat_gather and out.is_contiguous():
            # flattened views: gloo accepts only the concatenated 1-D form, NCCL accepts both
            self._dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1))
        else:
            parts = [torch_empty_like(t) for _ in range(self.world)]
            self._dist.all_gather(parts, t.contiguous())
            for r in range(self.world):
                out[r].copy_(parts[r])
        return out

    def gather_records(self, rec_u8, dtype: np.dtype) -> np.ndarray:
        """All ranks' copies of a device-resident record (uint8 tensor of dtype.itemsize bytes)."""
        raw = self.all_gather_tensor(rec_u8)
        return np.ascontiguousarray(raw).view(dtype).reshape(-1)

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


def reduce_safe_records(recs) -> dict:
    """Combine the per-rank records of so_sets_reduce_safe (structured array, one entry per rank) into the
    global one (gp_opt.py:504, :512, :634-636, :708-712); ties go to the lowest global row."""
    if len(recs) == 1:          # single rank: nothing to combine (this runs once per optimize(), which is host-bound on small grids)
        r = recs[0]
        return dict(n_safe=int(r["n_safe"]), max_l0=float(r["max_l0"]), argmax_l0=int(r["argmax_l0"]), max_u0=float(r["max_u0"]),
                    argmax_u0=int(r["argmax_u0"]))
    best_l, row_l = combine_max_first(recs["max_l0"], recs["argmax_l0"])
    best_u, row_u = combine_max_first(recs["max_u0"], recs["argmax_u0"])
    return dict(n_safe=int(np.sum(recs["n_safe"])), max_l0=best_l, argmax_l0=row_l, max_u0=best_u, argmax_u0=row_u)


def reduce_max_records(recs, scaling0) -> dict:
    """Combine the per-rank records of so_sets_maximizers (gp_opt.py:511-513, :642-644)."""
    if len(recs) == 1:
        r = recs[0]
        return dict(n_max=int(r["n_max"]), max_var=float(r["max_width0"]) / scaling0, best_value=float(r["best_value"]),
                    best_row=int(r["best_row"]))
    value, row = combine_max_first(recs["best_value"], recs["best_row"])
    return dict(n_max=int(np.sum(recs["n_max"])), max_var=float(np.max(recs["max_width0"])) / scaling0, best_value=value,
                best_row=row)


def gather_ragged(comm: Comm, arr: np.ndarray) -> np.ndarray:
    """Concatenate per-rank 1-D arrays of different lengths in rank order."""
    if not comm.active:
        return arr
    counts = comm.all_gather(np.array([arr.shape[0]], dtype=np.int64)).ravel()
    cap = max(int(counts.max()), 1)
    pad = np.zeros(cap, dtype=arr.dtype)
    pad[:arr.shape[0]] = arr
    allp = comm.all_gather(pad)
    return np.concatenate([allp[r, :counts[r]] for r in range(comm.world)])


def gather_row_blocks(comm: Comm, local: np.ndarray, n_rows: int) -> np.ndarray:
    """Reassemble a row-sharded array (blocks from shard_bounds) on every rank."""
    if not comm.active:
        return local
    per = -(-n_rows // comm.world)
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    allr = comm.all_gather(pad)
    return allr.reshape((-1,) + local.shape[1:])[:n_rows]


def gather_padded_rows(comm: Comm, local, total: int):
    """Reassemble a row-sharded torch tensor (blocks from shard_bounds) on every rank without leaving the
    device: blocks are padded to ceil(total/world) rows, all-gathered with one collective and trimmed.
    Works for CUDA tensors over NCCL and CPU tensors over gloo."""
    if not comm.active:
        return local
    import torch
    per = -(-int(total) // comm.world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((comm.world,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
    comm.all_gather_into(out, pad)
    return out.reshape((-1,) + tuple(local.shape[1:]))[:total]


def shard_stacked_blocks(full: np.ndarray, blocks: int, total: int, lo: int, hi: int) -> np.ndarray:
    """``full`` is ``blocks`` stacked (total, d) arrays (how the reference draws its PSO randoms,
    swarm.py:105: ``rand(2 * swarm_size, ndim)``); returns rows [lo, hi) of each block, stacked."""
    return np.concatenate([full[b * total + lo: b * total + hi] for b in range(blocks)], axis=0)


def order_candidates(comm: Comm, rows_sorted: np.ndarray, keys_sorted: np.ndarray) -> np.ndarray:
    """Global visiting order of expander candidates: key descending, ties by DESCENDING row -- what the
    reference's ``argsort()[::-1]`` (gp_opt.py:545-552) yields wherever NumPy's sort is stable."""
    if not comm.active:
        return rows_sorted
    allr, allk = gather_ragged(comm, rows_sorted), gather_ragged(comm, keys_sorted)
    return allr[np.lexsort((-allr, -allk))]
