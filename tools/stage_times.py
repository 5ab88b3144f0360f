#!/usr/bin/env python
"""Per-kernel times of one pass at a given batch (CUDA events between kernels; includes launch gaps)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="L")
ap.add_argument("--batches", default="512,2048,8192,16384")
args = ap.parse_args()
I, H, nh, O = synth.SHAPES[args.shape]
dnn = qd.QuantizedDnn.load_from_file(synth.network_file(args.shape), device=0)
for m in [int(x) for x in args.batches.split(",")]:
    d_in = torch.from_numpy(synth.make_frames(m, I, seed=7)).cuda()
    d_out = torch.empty(m, O, dtype=torch.float32, device="cuda")
    ctx = dnn.get_new_lazy_context(m)
    ms = ctx.profile_stages(d_in.data_ptr(), m, d_out.data_ptr(), iters=10)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        ctx.forward_device(d_in.data_ptr(), m, d_out.data_ptr(), s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.forward_device(d_in.data_ptr(), m, d_out.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 10
    hid = float(np.mean(ms[1:-2]))
    print(f"batch {m}: step {step*1e3:.1f} us = {m/step*1e3:.0f} frames/s | input {ms[0]*1e3:.1f} us, hidden avg {hid*1e3:.1f} us "
          f"({2*m*H*H/hid/1e9:.0f} TOP/s), output {ms[-2]*1e3:.1f} us ({2*m*H*O/ms[-2]/1e9:.0f} TOP/s), softmax {ms[-1]*1e3:.1f} us "
          f"({m*O*8/ms[-1]/1e6:.0f} GB/s)")
    ctx.delete()
    del d_in, d_out
dnn.delete()
