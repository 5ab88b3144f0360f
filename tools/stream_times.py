#!/usr/bin/env python
"""Stage times of a long-batch pass (layer-by-layer kernels) and whole-pass rate.  python tools/stream_times.py [frames]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=0)
x = torch.from_numpy(synth.make_frames(m, 440, seed=3)).to(dev)
y = torch.empty(m, 8000, dtype=torch.float32, device=dev)
ctx = dnn.get_new_lazy_context(m)
ms = ctx.profile_stages(x.data_ptr(), m, y.data_ptr(), iters=5)
ms = ctx.profile_stages(x.data_ptr(), m, y.data_ptr(), iters=10)
nl = dnn.layer_count()
print(f"M={m} stages us: input {ms[0]*1e3:.1f} hidden {[round(float(v)*1e3,1) for v in ms[1:nl-1]]} output {ms[nl-1]*1e3:.1f} softmax {ms[nl]*1e3:.1f} sum {ms.sum()*1e3:.1f}")
s = torch.cuda.current_stream()
for _ in range(4):
    ctx.forward_device(x.data_ptr(), m, y.data_ptr(), s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(30):
    ctx.forward_device(x.data_ptr(), m, y.data_ptr(), s.cuda_stream)
e1.record(s)
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 30
print(f"M={m} whole pass {t*1e3:.1f} us  -> {m/t*1e3/1e6:.2f} M frames/s   (FDNN_OUTPUT_SUB_ROWS={os.environ.get('FDNN_OUTPUT_SUB_ROWS','default')})")
ctx.delete()
dnn.delete()
