#!/usr/bin/env python
"""What the box can move from the GPUs into page-locked host memory: the ceiling of the end-to-end figure, which returns
32 000 bytes of scores per frame (8000 fp32 outputs) through PCIe.

  python tools/d2h_ceiling.py                         one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/d2h_ceiling.py

Each rank copies 16 MB (one 512-frame result) device → pinned host, `depth` copies in flight on two streams, for about
half a second; all ranks run at the same time (barrier).  Prints one JSON line (rank 0): per-rank and aggregate GB/s, and
the frames/s the aggregate corresponds to.  bench.py imports measure() for its e2e leg.
"""
from __future__ import annotations

import json
import os
import time


def measure(torch, dev, seconds: float = 0.5, nbytes: int = 512 * 8000 * 4, depth: int = 4, h2d_bytes: int = 0):
    """GB/s of `depth` outstanding D2H copies of `nbytes` (plus an optional H2D of h2d_bytes each) on device `dev`"""
    src = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(depth)]
    dst = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    up_src = [torch.empty(max(h2d_bytes, 1), dtype=torch.uint8).pin_memory() for _ in range(depth)]
    up_dst = [torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device=dev) for _ in range(depth)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def run(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % 2]):
                if h2d_bytes:
                    up_dst[i % depth].copy_(up_src[i % depth], non_blocking=True)
                dst[i % depth].copy_(src[i % depth], non_blocking=True)

    run(2 * depth)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    run(8)
    torch.cuda.synchronize(dev)
    per = (time.perf_counter() - t0) / 8
    n = max(16, int(seconds / max(per, 1e-6)))
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    run(n)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return n * nbytes / dt / 1e9, dt


def main():
    import torch
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    gbs, dt = measure(torch, dev)
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        all_ = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(all_, t)
        per_rank = [float(x.item()) for x in all_]
        dist.destroy_process_group()
    else:
        per_rank = [gbs]
    if rank == 0:
        total = sum(per_rank)
        print(json.dumps({"d2h_pinned_gbs_per_rank": per_rank, "d2h_pinned_gbs_total": total, "n_gpus": world,
                          "frames_per_s_at_32000_bytes": total * 1e9 / 32000.0, "copy_bytes": 512 * 8000 * 4,
                          "host_cores": len(os.sched_getaffinity(0))}))


if __name__ == "__main__":
    main()
