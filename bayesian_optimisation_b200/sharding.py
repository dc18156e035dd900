"""Multi-GPU sweep, one process per GPU (torchrun): candidates sharded by contiguous flat-index ranges,
the Cholesky replicated, and ONE exchange of 24-byte (score, index, nan) records to pick the winner.

The reference has no distributed code; this is the sharding its north star
describes.  Because shards are contiguous and ordered by rank, "smallest flat index among exact
ties" (point_selector.py:207) is preserved by reducing with (largest score, smallest index).
NCCL has no MAXLOC, so the exchange is an all_gather of the records followed by the same
deterministic reduce on every rank:

  * device path (NCCL): the sweep leaves its record in HBM (`GPEngine.acquire(..., sync=False)`),
    `all_gather_into_tensor` reads it from there, `bogp_reduce_results` folds the gathered records on
    the device, and the host reads 24 bytes once at the very end -- no host round trip per rank;
  * host path (gloo, CPU tests): the same records as bytes.

torch is imported when a function here is called, not when the package is imported (the single-process
path of `PointSelector` never needs it).
"""
from __future__ import annotations

import struct
from typing import Iterable, List, Sequence, Tuple

import numpy as np

NO_INDEX = (1 << 63) - 1
_REC = struct.Struct("<dqii")        # struct bogp_result (include/bogp.h)


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


def _dist():
    import torch.distributed as dist
    return dist


def _active(group=None) -> bool:
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def _collective_device(group=None):
    """Where collective buffers must live: the current CUDA device for NCCL, the host for gloo."""
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if _dist().get_backend(group) == "nccl" else torch.device("cpu")


def allreduce_maxloc(score: float, index: int, nan_flag: bool = False, device=None, group=None) -> Tuple[float, int]:
    """One collective on host-held values: all_gather of one 24-byte record per rank, then `reduce_pairs` on every
    rank.  A NaN flag raised by ANY rank raises IndexError on EVERY rank (the reference's behaviour on a NaN
    acquisition value, point_selector.py:207)."""
    if not _active(group):
        if nan_flag:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0 (NaN acquisition value)")
        return float(score), int(index)
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    raw = _REC.pack(float(score), int(index), 1 if nan_flag else 0, 0)
    mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone().to(_collective_device(group) if device is None else device)
    out = torch.empty(_REC.size * world, dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    blob = out.cpu().numpy().tobytes()
    recs = [_REC.unpack_from(blob, _REC.size * r) for r in range(world)]
    if any(r[2] for r in recs):
        raise IndexError("index 0 is out of bounds for axis 0 with size 0 (NaN acquisition value on some rank)")
    return reduce_pairs((r[0], r[1]) for r in recs)


def allreduce_maxloc_device(engine, record, group=None) -> Tuple[float, int]:
    """The device path: `record` is the 24-byte uint8 CUDA tensor a sweep left behind (`AcquireResult.record`).
    all_gather straight from it, fold on the device (bogp_reduce_results), ONE 24-byte host read."""
    if not _active(group):
        return engine.reduce_records(record, 1)
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    out = torch.empty(_REC.size * world, dtype=torch.uint8, device=record.device)
    dist.all_gather_into_tensor(out, record, group=group)
    return engine.reduce_records(out, world)


def allreduce_minloc(value: float, index: int, nan_flag: bool = False, device=None, group=None) -> Tuple[float, int]:
    """(smallest value, then smallest index) across ranks -- the restart that wins a sharded
    length-scale fit.  Same exchange as `allreduce_maxloc`."""
    s, i = allreduce_maxloc(-float(value), index, nan_flag=nan_flag or value != value, device=device, group=group)
    return -s, i


def all_gather_slices(parts: Sequence[np.ndarray], c_total: int, rank: int, world: int, group=None) -> List[np.ndarray]:
    """Every rank contributes its `shard_range` slice of each array in `parts` (float64); every rank gets the full
    length-`c_total` arrays back.  ONE all_gather for all arrays (slices are padded to the common ceil size)."""
    import torch
    dist = _dist()
    per = -(-c_total // world)
    k = len(parts)
    dev = _collective_device(group)
    mine = torch.zeros((k, per), dtype=torch.float64)
    for j, p in enumerate(parts):
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        mine[j, :len(p)] = torch.from_numpy(p)
    mine = mine.reshape(-1).to(dev)
    out = torch.empty(world * k * per, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    full = out.reshape(world, k, per).permute(1, 0, 2).reshape(k, world * per)[:, :c_total].cpu().numpy()
    return [np.ascontiguousarray(full[j]) for j in range(k)]


def all_gather_strided(part: np.ndarray, r_total: int, rank: int, world: int, group=None) -> np.ndarray:
    """Restart tables: rank r holds entries r, r+G, ...; every rank gets the full table back."""
    import torch
    dist = _dist()
    per = -(-r_total // world)
    dev = _collective_device(group)
    mine = torch.zeros(per, dtype=torch.float64)
    mine[:len(part)] = torch.from_numpy(np.asarray(part, dtype=np.float64).reshape(-1))
    mine = mine.to(dev)
    out = torch.empty(world * per, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return np.ascontiguousarray(out.cpu().numpy().reshape(world, per).T.reshape(-1)[:r_total])


def sharded_nlml_argmin(engine, x, y, ells, rank: int, world: int, jitter=None, group=None):
    """Multi-restart length-scale selection across GPUs: rank r evaluates restarts r, r+G, ... with one
    batched launch (K3), rounds to float32 like the reference's table (point_selector.py:126) and the
    ranks agree on the first minimum (lowest restart id among ties, point_selector.py:141).
    A NaN anywhere in the GLOBAL table raises IndexError on every rank -- what the reference's
    `np.argwhere(nlogml == np.amin(nlogml))[0]` does (amin is NaN, nothing compares equal) -- whatever
    the number of ranks.  Returns (nlml_float32, restart_id, local_table)."""
    ids = np.arange(rank, len(ells), world)
    kw = {} if jitter is None else {"jitter": jitter}
    nan = False
    if len(ids):
        table = engine.nlml_batched(x, y, np.asarray(ells)[ids], **kw)
        table = (table.cpu().numpy() if hasattr(table, "cpu") else np.asarray(table)).astype(np.float32)
        nan = bool(np.isnan(table).any())
        k = 0 if nan else int(np.flatnonzero(table == np.amin(table))[0])
        val, idx = (0.0 if nan else float(table[k])), int(ids[k])
    else:
        table, val, idx = np.zeros(0, np.float32), float("inf"), NO_INDEX
    gv, gi = allreduce_minloc(val, idx, nan_flag=nan, group=group)
    return gv, gi, table


def sharded_acquire(engine, fit, candidates, c_total: int, rank: int, world: int, group=None, **kw):
    """Score this rank's slice on its GPU and reduce on the device.  Returns (score, index, local AcquireResult)."""
    b, e = shard_range(c_total, rank, world)
    if e > b:
        engine.set_global_seed(world > 1)      # screened arg-max-only sweeps: every shard screens against the same floor
        try:
            res = engine.acquire(fit, candidates, b, e, sync=False, **kw)
        finally:
            engine.set_global_seed(False)
        rec = res.record
    else:
        res, rec = None, engine.empty_record()
    gs, gi = allreduce_maxloc_device(engine, rec, group=group)
    return gs, gi, res
