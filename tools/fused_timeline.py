#!/usr/bin/env python
"""Per-layer phase latencies inside the fused kernel (nanosecond stamps, fdnn_ctx_timeline).  python tools/fused_timeline.py [L|S] [frames]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "L"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
dnn = qd.QuantizedDnn.load_from_file(synth.network_file(shape), device=0)
dnn.set_tile_policy(os.environ.get("TIMELINE_POLICY", "latency"))
I, O = dnn.input_dimension(), dnn.output_dimension()
x = torch.from_numpy(synth.make_frames(m, I, seed=3)).to(dev)
y = torch.empty(m, O, dtype=torch.float32, device=dev)
ctx = dnn.get_new_lazy_context(m)
for _ in range(3):
    ctx.forward_device(x.data_ptr(), m, y.data_ptr())
torch.cuda.synchronize()
ctx.timeline(True)
ctx.forward_device(x.data_ptr(), m, y.data_ptr())
torch.cuda.synchronize()
t = ctx.timeline(False).astype(np.int64)  # [layers][1024][8]
nl = dnn.layer_count() - 1
used = t[0, :, 1] != 0
n_cta = int(used.sum())
t0 = t[0, used, 7].min()
print(f"{shape} M={m}: {n_cta} CTAs; times in us relative to the first producer entering layer 0")
names = ["rowblock ready", "first stage landed", "accumulator ready", "scan done", "stores issued", "released", "weights requested", "producer enters"]
for j in range(nl):
    tj = (t[j, used, :] - t0) / 1e3
    order = [7, 6, 0, 1, 2, 3, 4, 5]
    print(f"layer {j}: " + "  ".join(f"{names[k]} {np.median(tj[:, k]):7.2f} (max {tj[:, k].max():7.2f})" for k in order))
ctx.delete()
dnn.delete()
