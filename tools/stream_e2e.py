#!/usr/bin/env python
"""End-to-end rate of a frame stream through QuantizedDnn.calculate on pinned buffers, per rank (torchrun) — an experiment
harness for the configs[4] leg of bench.py.  Args: --call FRAMES --threads T --frames TOTAL"""
import argparse
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import quantized_dnn as qd, sharding, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--call", type=int, default=16384)
ap.add_argument("--threads", type=int, default=2)
ap.add_argument("--frames", type=int, default=500_000, help="frames per rank")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=local)
h_in = [qd.PinnedArray((args.call, 440), np.float32) for _ in range(args.threads)]
h_out = [qd.PinnedArray((args.call, 8000), np.float32) for _ in range(args.threads)]
for j in range(args.threads):
    h_in[j].array[:] = synth.make_frames(args.call, 440, seed=400 + rank * 10 + j)
    dnn.calculate(h_in[j].array, out=h_out[j].array)
    dnn.calculate(h_in[j].array, out=h_out[j].array)
spans = list(sharding.chunk_ranges(0, args.frames, args.call))
busy = [0.0] * args.threads


def worker(t_):
    torch.cuda.set_device(local)
    for k in range(t_, len(spans), args.threads):
        a, b = spans[k]
        t1 = time.perf_counter()
        dnn.calculate(h_in[t_].array[: b - a], out=h_out[t_].array[: b - a])
        busy[t_] += time.perf_counter() - t1


torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
ws = [threading.Thread(target=worker, args=(t_,)) for t_ in range(args.threads)]
[w.start() for w in ws]
[w.join() for w in ws]
secs = time.perf_counter() - t0
t = torch.tensor([secs], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world} call {args.call} threads {args.threads} sync {os.environ.get('FDNN_SYNC', 'sleep')}: {world * args.frames / float(t.item()) / 1e6:.3f} M frames/s "
          f"({world * args.frames * 32000 / float(t.item()) / 1e9:.1f} GB/s down); mean call {1e3 * sum(busy) / len(spans):.2f} ms", flush=True)
dnn.delete()
if world > 1:
    dist.destroy_process_group()
