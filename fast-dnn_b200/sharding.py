"""Multi-GPU plumbing for the one thing the path needs: frames are independent, so
a frame stream is cut into contiguous shards, one per rank, every rank holds a full replica of
the packed weights delivered by ONE broadcast at load, and there is no per-frame collective.

One process per GPU, launched by torchrun; torch.distributed is only the transport (NCCL on GPUs,
gloo in the CPU tests).  Nothing here computes.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of rank's frames; sizes differ by at most one, order preserved."""
    if world <= 0 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def chunk_ranges(begin: int, end: int, chunk: int):
    """[begin, end) cut into pieces of at most `chunk` frames (the streaming unit of one GPU)."""
    if chunk <= 0:
        raise ValueError("chunk must be positive")
    for lo in range(begin, end, chunk):
        yield lo, min(lo + chunk, end)


def broadcast_blob(blob, src: int = 0, device=None):
    """Rank `src` passes the packed model (numpy uint8 from quantized_dnn.pack); everyone gets a
    torch uint8 tensor holding the same bytes on `device` (cuda → feed data_ptr() to
    QuantizedDnn.load_from_blob).  Two collectives: the size, then the payload."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    device = torch.device(device) if device is not None else torch.device("cpu")
    if rank == src:
        if blob is None:
            raise ValueError("source rank must provide the blob")
        payload = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8)).to(device)
        size = torch.tensor([payload.numel()], dtype=torch.int64, device=device)
    else:
        size = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(size, src)
    if rank != src:
        payload = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(payload, src)
    return payload


def gather_counts(local_count: int, device=None) -> list[int]:
    """Frames processed per rank (for throughput = Σ frames ÷ max-over-ranks time)."""
    import torch
    import torch.distributed as dist

    device = torch.device(device) if device is not None else torch.device("cpu")
    mine = torch.tensor([local_count], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [int(t.item()) for t in out]
