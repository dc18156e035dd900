"""Multi-GPU sweep: one process per GPU, candidates sharded by contiguous flat-index ranges,
the Cholesky replicated, and ONE 16-byte-per-rank exchange to pick the winner.

The reference has no distributed code; this is the sharding its north star
describes.  Because shards are contiguous and ordered by rank, "smallest flat index among exact
ties" (point_selector.py:207) is preserved by reducing with (largest score, smallest index).
NCCL has no MAXLOC, so the exchange is an all_gather of (score, index) pairs followed by the
same deterministic reduce on every rank.
"""
from __future__ import annotations

import struct
from typing import Iterable, Tuple

import torch
import torch.distributed as dist

NO_INDEX = (1 << 63) - 1


def shard_range(c_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Flat indices [begin, end) scored by `rank`: ceil(C/G)-sized contiguous slices."""
    per = -(-c_total // world)
    b = min(c_total, rank * per)
    return b, min(c_total, b + per)


def restart_slice(r_total: int, rank: int, world: int) -> range:
    """Restarts handled by `rank`: r, r+G, r+2G, ...."""
    return range(rank, r_total, world)


def reduce_pairs(pairs: Iterable[Tuple[float, int]]) -> Tuple[float, int]:
    """(largest score, then smallest index); NaN scores never win.  Pure function."""
    best_s, best_i = float("-inf"), NO_INDEX
    for s, i in pairs:
        if s != s:
            continue
        if s > best_s or (s == best_s and i < best_i):
            best_s, best_i = s, i
    return best_s, best_i


def allreduce_maxloc(score: float, index: int, device=None, group=None) -> Tuple[float, int]:
    """One collective: all_gather of 16 bytes per rank, then `reduce_pairs` on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(score), int(index)
    world = dist.get_world_size(group)
    raw = struct.pack("<dq", float(score), int(index))
    mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    out = torch.empty(16 * world, dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    blob = bytes(out.cpu().numpy().tobytes())
    pairs = [struct.unpack_from("<dq", blob, 16 * r) for r in range(world)]
    return reduce_pairs(pairs)


def allreduce_minloc(value: float, index: int, device=None, group=None) -> Tuple[float, int]:
    """(smallest value, then smallest index) across ranks -- the restart that wins a sharded
    length-scale fit.  Same 16-byte exchange as `allreduce_maxloc`."""
    s, i = allreduce_maxloc(-float(value) if value == value else float("nan"), index, device=device, group=group)
    return -s, i


def sharded_nlml_argmin(engine, x, y, ells, rank: int, world: int, jitter=None, group=None):
    """Multi-restart length-scale selection across GPUs: rank r evaluates restarts r, r+G, ... with one
    batched launch (K3), rounds to float32 like the reference's table (point_selector.py:126) and the
    ranks agree on the first minimum (lowest restart id among ties, point_selector.py:141).
    Returns (nlml_float32, restart_id, local_table)."""
    import numpy as np
    ids = np.arange(rank, len(ells), world)
    kw = {} if jitter is None else {"jitter": jitter}
    if len(ids):
        table = engine.nlml_batched(x, y, np.asarray(ells)[ids], **kw).cpu().numpy().astype(np.float32)
        k = int(np.flatnonzero(table == np.amin(table))[0]) if not np.isnan(table).any() else 0
        val, idx = float(table[k]), int(ids[k])
        if np.isnan(table).any():
            val = float("nan")
    else:
        table, val, idx = np.zeros(0, np.float32), float("inf"), NO_INDEX
    use_dev = dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
    gv, gi = allreduce_minloc(val, idx, device=engine.device if use_dev else None, group=group)
    return gv, gi, table


def sharded_acquire(engine, fit, candidates, c_total: int, rank: int, world: int, group=None, **kw):
    """Score this rank's slice on its GPU and reduce.  Returns (score, index, local AcquireResult)."""
    b, e = shard_range(c_total, rank, world)
    if e > b:
        res = engine.acquire(fit, candidates, b, e, **kw)
        s, i = res.best_score, res.best_index
    else:
        res, s, i = None, float("-inf"), NO_INDEX
    gs, gi = allreduce_maxloc(s, i, device=engine.device if dist.is_initialized() and dist.get_backend(group) == "nccl" else None, group=group)
    return gs, gi, res
