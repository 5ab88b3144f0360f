#!/usr/bin/env python
"""Prints per-phase SM-clock timings of the tensor-core layer kernels for one forward pass
(profiling aid; see fdnn_ctx_timeline in include/fdnn.h)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="L")
ap.add_argument("--batch", type=int, default=512)
args = ap.parse_args()
I, H, nh, O = synth.SHAPES[args.shape]
dnn = qd.QuantizedDnn.load_from_file(synth.network_file(args.shape), device=0)
d_in = torch.from_numpy(synth.make_frames(args.batch, I, seed=7)).cuda()
d_out = torch.empty(args.batch, O, dtype=torch.float32, device="cuda")
ctx = dnn.get_new_lazy_context(args.batch)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.forward_device(d_in.data_ptr(), args.batch, d_out.data_ptr(), s)
torch.cuda.synchronize()
ctx.timeline(True)
ctx.forward_device(d_in.data_ptr(), args.batch, d_out.data_ptr(), s)
torch.cuda.synchronize()
tl = ctx.timeline(False).astype(np.int64)
names = ["setup", "first operands", "mma issue done", "scan done", "epilogue sees acc", "tile done", "epilogue sees scan"]
for layer in range(tl.shape[0]):
    t = tl[layer]
    used = t[:, 0] > 0
    if not used.any():
        continue
    t = t[used]
    rel = t - t[:, :1]
    print(f"layer {layer}: {used.sum()} CTAs; median cycles since CTA entry:")
    for slot in range(1, 8):
        col = rel[:, slot][t[:, slot] > 0]
        if col.size:
            print(f"   {names[slot - 1]:<22} median {int(np.median(col)):>7}  min {int(col.min()):>7}  max {int(col.max()):>7}")
ctx.delete()
dnn.delete()
