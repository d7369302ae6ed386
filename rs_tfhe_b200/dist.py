"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink)
for the single collective the path has -- broadcasting the re-laid-out cloud key
once.  Gate batches are sharded by contiguous index range with no per-gate
communication (the reference's only parallelism is par_map over independent
ciphertexts, src/trgsw.rs:297-305)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(count: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `count` items for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return rank * count // world, (rank + 1) * count // world


class _DevBlob:
    """Exposes a raw device allocation through __cuda_array_interface__ so torch can
    wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_cloud_key(engine, cloud_key, src: int = 0, group=None) -> float:
    """Rank `src` uploads `cloud_key` (reference layout -> device layout); every other
    rank allocates the device blob and receives it by one NCCL broadcast.  Returns the
    broadcast time in ms (device events).  `cloud_key` may be None on non-src ranks, but
    every rank must pass the decomposition offset via `cloud_key_offset` semantics: the
    offset travels in a 1-element tensor alongside the blob."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    dev = torch.device("cuda", engine.device)
    if rank == src:
        if cloud_key is None:
            raise ValueError("source rank needs the cloud key")
        engine.load_cloud_key(cloud_key)
        off = torch.tensor([int(cloud_key.decomposition_offset)], dtype=torch.int64, device=dev)
    else:
        engine.alloc_cloud_key()
        off = torch.zeros(1, dtype=torch.int64, device=dev)
    ptr, nbytes = engine.cloud_key_blob()
    blob = torch.as_tensor(_DevBlob(ptr, nbytes), device=dev)
    torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    dist.broadcast(blob, src=src, group=group)
    dist.broadcast(off, src=src, group=group)
    t1.record()
    torch.cuda.synchronize(dev)
    if rank != src:
        engine.commit_cloud_key(int(off.item()) & 0xFFFFFFFF)
    return float(t0.elapsed_time(t1))


def gather_outputs(local_out: np.ndarray, count: int, world: int, group=None) -> Optional[np.ndarray]:
    """Host-side gather of per-rank output shards to rank 0 (order preserved)."""
    import torch.distributed as dist

    rank = dist.get_rank(group)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local_out, parts, dst=0, group=group)
    if rank != 0:
        return None
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == count
    return out
