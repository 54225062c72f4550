"""Host-side helpers for the row-sharded multi-GPU path.

One process per GPU (``torchrun``); ``torch.distributed`` is used only to
bootstrap: the NCCL unique id and the shard sizes travel over it (any backend,
``gloo`` works), while every data-path collective runs inside the library on
its own NCCL communicator and stream.
"""

from __future__ import annotations

import numpy as np


def is_distributed() -> bool:
    try:
        import torch.distributed as td
    except Exception:  # pragma: no cover
        return False
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def world():
    """``(rank, world_size)`` of the current torch.distributed job (0, 1 if none)."""
    if not is_distributed():
        return 0, 1
    import torch.distributed as td
    return td.get_rank(), td.get_world_size()


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Broadcast a small byte string from ``src`` (used for the NCCL unique id)."""
    import torch.distributed as td
    box = [payload]
    td.broadcast_object_list(box, src=src)
    return box[0]


def allgather_ints(value: int) -> list[int]:
    import torch.distributed as td
    out = [None] * td.get_world_size()
    td.all_gather_object(out, int(value))
    return [int(v) for v in out]


def allgather_objects(obj) -> list:
    """Every rank's (picklable, small) object, in rank order."""
    import torch.distributed as td
    out = [None] * td.get_world_size()
    td.all_gather_object(out, obj)
    return out


def allreduce_array(arr: np.ndarray, op: str = "sum") -> np.ndarray:
    """Element-wise all-reduce of a small host array over the torch.distributed world (identity
    without one).  Used for host-side bookkeeping only (chunk sums of the Nystrom normalisation,
    document frequencies in streamed mode) -- the data path reduces inside the library over NCCL."""
    if not is_distributed():
        return arr
    import torch
    import torch.distributed as td
    rop = {"sum": td.ReduceOp.SUM, "min": td.ReduceOp.MIN, "max": td.ReduceOp.MAX}[op]
    t = torch.from_numpy(np.ascontiguousarray(arr).copy())
    try:
        td.all_reduce(t, op=rop)
    except RuntimeError:            # NCCL-only process group: reduce on the device
        t = t.cuda()
        td.all_reduce(t, op=rop)
        t = t.cpu()
    return t.numpy()


def chunk_overlaps(row0: int, n_local: int, chunk_size: int):
    """Global ``chunk_size`` row blocks that overlap the shard ``[row0, row0 + n_local)``:
    yields ``(chunk index, local start, local stop)``."""
    if n_local <= 0:
        return
    c = row0 // chunk_size
    while c * chunk_size < row0 + n_local:
        lo = max(c * chunk_size, row0) - row0
        hi = min((c + 1) * chunk_size, row0 + n_local) - row0
        yield c, lo, hi
        c += 1


def shard_offsets(n_locals: list[int]) -> list[int]:
    """Global row offset of every rank's shard from the list of shard sizes."""
    offs, run = [], 0
    for v in n_locals:
        offs.append(run)
        run += int(v)
    return offs


def balanced_row_splits(indptr: np.ndarray, parts: int) -> np.ndarray:
    """Contiguous row blocks balanced by stored entries.

    Returns ``parts + 1`` row boundaries; block ``p`` is rows
    ``[b[p], b[p+1])`` and holds ~nnz/parts entries.
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.shape[0] - 1
    nnz = int(indptr[-1])
    targets = (np.arange(1, parts, dtype=np.float64) * nnz / parts)
    cuts = np.searchsorted(indptr, targets, side="left")
    bounds = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def equal_row_splits(n: int, parts: int) -> np.ndarray:
    """``parts + 1`` boundaries of near-equal contiguous row blocks."""
    return (np.arange(parts + 1, dtype=np.int64) * n) // parts


def attach_engine_comm(engine) -> None:
    """Join ``engine`` to a library-level NCCL communicator spanning the
    current torch.distributed world (rank 0 mints the id, everyone gets it)."""
    rank, ws = world()
    if ws == 1:
        return
    uid = engine.unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, src=0)
    engine.init_comm(rank, ws, uid)
