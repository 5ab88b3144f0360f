import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from fast_dnn_b200 import quantized_dnn as qd, synth
dnn = qd.QuantizedDnn.load_from_file(synth.network_file("S"))
x = synth.make_frames(300, 440, seed=4)
out = dnn.calculate(x)
print("rows sum", float(out.sum(axis=1).mean()))
dnn.delete()
