#!/usr/bin/env python
"""Pass times with and without the fused int8-stack kernel (csrc/qlayer_fused.cu), per batch size and tile policy.
Run on the GPU box:  python tools/fused_times.py [L|S]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "L"
dev = torch.device("cuda", 0)
dnn = qd.QuantizedDnn.load_from_file(synth.network_file(shape), device=0)
I, O = dnn.input_dimension(), dnn.output_dimension()
for policy in ("latency", "throughput"):
    dnn.set_tile_policy(policy)
    for m in (128, 512, 1024):
        x = torch.from_numpy(synth.make_frames(m, I, seed=3)).to(dev)
        y = torch.empty(m, O, dtype=torch.float32, device=dev)
        ctx = dnn.get_new_lazy_context(m)
        row = []
        for flag in ("2", "0"):
            os.environ["FDNN_FUSED"] = flag
            ctx.profile_pass(x.data_ptr(), m, y.data_ptr(), iters=5)
            t_in, t_rest, fused = ctx.profile_pass(x.data_ptr(), m, y.data_ptr(), iters=50)
            row.append((fused, t_in * 1e3, t_rest * 1e3))
        ctx.delete()
        print(f"{shape} policy={policy:10s} M={m:5d}  " + "   ".join(f"fused={f!s:5s} input {a:7.1f} us  rest {b:7.1f} us" for f, a, b in row), flush=True)
os.environ.pop("FDNN_FUSED", None)
dnn.delete()
