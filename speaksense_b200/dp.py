"""Data-parallel host logic: clips / streams are independent units, so the path shards
with no data-path collective.  Rank r of W owns a contiguous slice of the batch (256 clips -> 32 per
GPU in BASELINE config 4); the only exchange is at init, when rank 0's ncclUniqueId is ferried to the
other ranks so that the C++ engine can ncclBroadcast the packed weight arena."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def shard_bounds(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n_units for `rank` (first n_units % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(units: Sequence, rank: int, world: int) -> List:
    lo, hi = shard_bounds(len(units), rank, world)
    return list(units[lo:hi])


def broadcast_bytes(payload: Optional[bytes], src: int = 0) -> bytes:
    """Ferry a small byte string (the 128-byte ncclUniqueId) from `src` to every rank over the
    already-initialised torch.distributed group (gloo or nccl)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def gather_results(local: List, dst: int = 0) -> Optional[List]:
    """Transcripts are host strings: gather the per-rank result lists on `dst` in rank (= clip) order."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    if out is None:
        return None
    flat = []
    for part in out:
        flat.extend(part)
    return flat
