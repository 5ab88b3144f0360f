"""A small pass for compute-sanitizer (memcheck / racecheck / synccheck): 300 frames through the 440-4x512-2000 network, once
through the fused kernel (FDNN_FUSED=2) or layer by layer (FDNN_FUSED=0), plus a lazy row.  Usage (under gpurun):
  FDNN_FUSED=2 compute-sanitizer --tool memcheck python tools/sanitizer_case.py"""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np  # noqa: E402
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

dnn = qd.QuantizedDnn.load_from_file(synth.network_file("S"))
x = synth.make_frames(300, 440, seed=4)
out = dnn.calculate(x)
out = dnn.calculate(x)
print("rows sum", float(out.sum(axis=1).mean()))
ctx = dnn.get_new_lazy_context(300)
ctx.calculate_until_output(x)
row = ctx.calculate_for_output_nodes(np.ones(2000, np.int8))
print("lazy row sum", float(row.sum()))
ctx.delete()
dnn.delete()
