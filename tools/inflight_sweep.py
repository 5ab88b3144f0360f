#!/usr/bin/env python
"""Device-resident throughput at batch 512 with 1..8 contexts in flight on their own streams (run under gpurun).
bench.py uses 4: 144 -> 124 -> 115 -> 112 us per step for 1 -> 2 -> 3 -> 4 on B200."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_dnn_b200 import quantized_dnn as qd, synth
B, I, O = int(os.environ.get("B", "512")), 440, 8000
dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=0)
if os.environ.get("POLICY"):
    dnn.set_tile_policy(os.environ["POLICY"])
for nctx in [int(x) for x in os.environ.get("SWEEP", "1,2,3,4,6,8").split(",")]:
    pool = 16
    d_in = [torch.from_numpy(synth.make_frames(B, I, seed=100 + i)).cuda() for i in range(pool)]
    d_out = [torch.empty(B, O, dtype=torch.float32, device="cuda") for _ in range(pool)]
    ctxs = [dnn.get_new_lazy_context(B) for _ in range(nctx)]
    streams = [torch.cuda.Stream() for _ in range(nctx)]
    def step(i):
        c = i % nctx
        ctxs[c].forward_device(d_in[i % pool].data_ptr(), B, d_out[i % pool].data_ptr(), streams[c].cuda_stream)
    for i in range(32): step(i)
    torch.cuda.synchronize()
    K = 400
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = [torch.cuda.Event(enable_timing=True) for _ in range(nctx)]
    e0.record(torch.cuda.current_stream())
    for s in streams: s.wait_event(e0)
    for i in range(K): step(i)
    for s, e in zip(streams, e1): e.record(s)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e) for e in e1)
    print(f"{nctx} context(s)/stream(s): {ms/K*1e3:.1f} us per step, {B*K/ms*1e3:.0f} frames/s (wall {(time.perf_counter()-t0)/K*1e6:.1f} us/step)")
    for c in ctxs: c.delete()
dnn.delete()
