#!/usr/bin/env python
"""Device timeline of the batch-512 loop with several contexts in flight (CUPTI through torch.profiler): per-kernel start / end on
every stream, written as a compact JSON list to gpurun_out/.  python tools/inflight_trace.py [contexts] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from fast_dnn_b200 import quantized_dnn as qd, synth

nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B, I, O = 512, 440, 8000
dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=0)
dnn.set_tile_policy(os.environ.get("POLICY", "throughput"))
pool = 16
d_in = [torch.from_numpy(synth.make_frames(B, I, seed=100 + i)).cuda() for i in range(pool)]
d_out = [torch.empty(B, O, dtype=torch.float32, device="cuda") for _ in range(pool)]
ctxs = [dnn.get_new_lazy_context(B) for _ in range(nctx)]
streams = [torch.cuda.Stream() for _ in range(nctx)]


def step(i):
    c = i % nctx
    ctxs[c].forward_device(d_in[i % pool].data_ptr(), B, d_out[i % pool].data_ptr(), streams[c].cuda_stream)


for i in range(96):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/inflight_trace_{nctx}.json"
prof.export_chrome_trace(path + ".full")
ev = [e for e in json.load(open(path + ".full"))["traceEvents"] if e.get("cat") == "kernel"]
out = [{"name": e["name"][:60], "ts": e["ts"], "dur": e["dur"], "stream": e["args"].get("stream"), "grid": e["args"].get("grid"),
        "block": e["args"].get("block"), "smem": e["args"].get("shared memory")} for e in ev]
json.dump(out, open(path, "w"))
os.remove(path + ".full")
print(len(out), "kernel records ->", path)
for c in ctxs:
    c.delete()
dnn.delete()
