import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fast_dnn_b200 import quantized_dnn as qd, synth
for shape, n in (("S", 128), ("L", 512), ("L", 4096)):
    I, H, nh, O = synth.SHAPES[shape]
    dnn = qd.QuantizedDnn.load_from_file(synth.network_file(shape))
    ctx = dnn.get_new_lazy_context(n)
    ctx.calculate_until_output(synth.make_frames(n, I, seed=7))
    u = ctx.input_undecided()
    print(shape, n, "undecided", u, "of", n * H, f"= {100.0 * u / (n * H):.2f} %" if u is not None else "")
    ctx.delete(); dnn.delete()
