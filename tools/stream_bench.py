#!/usr/bin/env python
"""BASELINE configs[4]: a long synthetic frame stream (default 1 M frames) through the 7×2048/8000
network, sharded contiguously over the ranks (one process per GPU under torchrun; works with 1).
Frames of a shard are produced chunk by chunk from the seeded generator (a small pool of distinct
chunks is cycled so that the generator is not what gets timed; `--verify` checks one chunk per rank
against the direct call).  Reports end-to-end frames/s (pinned host buffers, H2D and D2H inside).

  python tools/stream_bench.py --frames 1000000
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/stream_bench.py
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from fast_dnn_b200 import quantized_dnn as qd, sharding, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1_000_000)
ap.add_argument("--chunk", type=int, default=8192)
ap.add_argument("--threads", type=int, default=3)
ap.add_argument("--verify", action="store_true")
args = ap.parse_args()

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
path = synth.network_file("L")
if world > 1:
    blob = sharding.broadcast_blob(qd.pack(path) if rank == 0 else None, src=0, device=dev)
    dnn = qd.QuantizedDnn.load_from_blob(blob.data_ptr(), device=local, size=blob.numel())
else:
    dnn = qd.QuantizedDnn.load_from_file(path, device=local)
I, O = dnn.input_dimension(), dnn.output_dimension()
begin, end = sharding.shard_range(args.frames, rank, world)
chunks = list(sharding.chunk_ranges(begin, end, args.chunk))

pool = 4
h_in = [[qd.PinnedArray((args.chunk, I), np.float32) for _ in range(pool)] for _ in range(args.threads)]
h_out = [[qd.PinnedArray((args.chunk, O), np.float32) for _ in range(2)] for _ in range(args.threads)]
for t in range(args.threads):
    for j in range(pool):
        lo = begin + ((t * pool + j) * args.chunk) % max(end - begin, 1)
        h_in[t][j].array[:] = synth.make_frames(args.chunk, I, seed=7, start=lo)


def worker(t, my_chunks):
    torch.cuda.set_device(local)
    for k, (lo, hi) in enumerate(my_chunks):
        n = hi - lo
        dnn.calculate(h_in[t][k % pool].array[:n], 10, out=h_out[t][k % 2].array[:n])


def run(chunk_list):
    parts = [chunk_list[t::args.threads] for t in range(args.threads)]
    ws = [threading.Thread(target=worker, args=(t, parts[t])) for t in range(args.threads)]
    [w.start() for w in ws]
    [w.join() for w in ws]
    torch.cuda.synchronize()


run(chunks[: 2 * args.threads])  # warm-up: contexts, graphs
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
run(chunks)
secs = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([secs], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())
ok = None
if args.verify:
    n = min(args.chunk, 512)
    a = dnn.calculate(h_in[0][0].array[:n])
    b = dnn.calculate(h_in[0][0].array[:n][::-1].copy())[::-1]
    ok = bool(np.array_equal(a, b) and np.allclose(a.sum(axis=1), 1.0, atol=5e-5))
if rank == 0:
    print(json.dumps({"metric": "frames/sec, 1M-frame stream (7x2048 hidden, 8000 out)", "value": args.frames / secs, "unit": "frames/s",
                      "n_gpus": world, "frames": args.frames, "chunk": args.chunk, "host_threads_per_gpu": args.threads, "seconds": secs,
                      "h2d_bytes": args.frames * I * 4, "d2h_bytes": args.frames * O * 4, "verified": ok}))
dnn.delete()
if world > 1:
    dist.destroy_process_group()
