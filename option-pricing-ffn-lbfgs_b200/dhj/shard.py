"""Multi-GPU sharding of the pricing / calibration path: one process per GPU, `torch.distributed`.

The path shards by parameter set (or calibration instance) with NO data-path collective: every unit is
independent.  Rank g owns the contiguous block [g*ceil(n/G), (g+1)*ceil(n/G)).  The only
communication is the gather of results (prices, losses, calibrated parameters) at the end — NCCL over
NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: [lo, hi) of rank `rank` out of `world`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n_items // world) if n_items else 0
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def history_shard(n_samples: int, path_len: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) sample range of rank `rank` when a counter-stream dataset of `n_samples` samples is split by whole
    histories of `path_len` samples (`shard_bounds` over histories; the last history may be partial)."""
    if path_len < 1:
        raise ValueError("path_len must be >= 1")
    n_paths = -(-n_samples // path_len) if n_samples else 0
    q_lo, q_hi = shard_bounds(n_paths, world, rank)
    return min(n_samples, q_lo * path_len), min(n_samples, q_hi * path_len)


def _dist():
    # importing torch costs seconds: only look for a process group if somebody already imported it
    import sys
    if "torch" not in sys.modules:
        return None
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return None
    return dist


def gather_rows(local: np.ndarray, n_total: int, group=None, device=None) -> np.ndarray:
    """All-gather the row blocks produced under `shard_bounds` back into the full [n_total, ...] array.

    With an NCCL group pass `device` (the rank's cuda device) — the payload then travels GPU to GPU over
    NVLink; with gloo leave it None.  Without an initialised process group the input is returned as is.
    """
    dist = _dist()
    local = np.ascontiguousarray(local)
    if dist is None or dist.get_world_size(group) == 1:
        return local
    import torch
    world = dist.get_world_size(group)
    per = -(-n_total // world) if n_total else 0
    row_shape = local.shape[1:]
    padded = np.zeros((per,) + row_shape, dtype=local.dtype)
    padded[:local.shape[0]] = local
    t = torch.from_numpy(padded)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * per,) + row_shape, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy()[:n_total]


def price_grid_sharded(price_fn, params, S0, *args, group=None, device=None, gather=True, **kwargs):
    """Price this rank's block of parameter sets with `price_fn` (e.g. `ctx.price_grid`) and, if `gather`,
    return the full [P, nT, nK] array on every rank (else only the local block and its bounds)."""
    dist = _dist()
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    params = np.asarray(params, dtype=np.float64).reshape(-1, 13)
    P = params.shape[0]
    lo, hi = shard_bounds(P, world, rank)
    S0 = np.asarray(S0, dtype=np.float64).reshape(-1)
    s0_local = S0 if S0.size == 1 else S0[lo:hi]
    local = price_fn(params[lo:hi], s0_local, *args, **kwargs)
    if not gather:
        return local, (lo, hi)
    return gather_rows(local, P, group, device)
