#!/usr/bin/env python
"""A short, clean run of the hot path for ncu: W warm-up steps + K steps of the device-resident
forward pass (BASELINE configs[2] by default), nothing else.  Usage (under gpurun):
  ncu --metrics gpu__time_duration.sum --clock-control none -s <launches to skip> -c <N> --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="L")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--policy", default="latency", help="tile policy: latency (one caller) or throughput (several contexts in flight)")
args = ap.parse_args()

I, H, nh, O = synth.SHAPES[args.shape]
dnn = qd.QuantizedDnn.load_from_file(synth.network_file(args.shape), device=0)
dnn.set_tile_policy(args.policy)
d_in = torch.from_numpy(synth.make_frames(args.batch, I, seed=7)).cuda()
d_out = torch.empty(args.batch, O, dtype=torch.float32, device="cuda")
ctx = dnn.get_new_lazy_context(args.batch)
stream = torch.cuda.current_stream()
for _ in range(args.warmup + args.steps):
    ctx.forward_device(d_in.data_ptr(), args.batch, d_out.data_ptr(), stream.cuda_stream)
torch.cuda.synchronize()
print("rows sum to", float(d_out.sum(dim=1).mean()))
ctx.delete()
dnn.delete()
