#!/usr/bin/env python
"""Java_suskun_nn_QuantizedDnn_calculate through the C JVM stand-in (tools/jni_harness.c) with 1 … 12 caller threads.
python tools/jni_threads.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

h = C.CDLL(os.path.join(ROOT, "tools", "libjni_harness.so"))
h.jni_harness_calculate.restype = C.c_double
h.jni_harness_calculate.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
B, I, O = 512, 440, 8000
frames = synth.make_frames(B, I, seed=4242)
first = np.zeros((B, O), np.float32)
path = synth.network_file("L")
for threads in (1, 2, 4, 6, 8, 10, 12, 16):
    iters = 40
    secs = h.jni_harness_calculate(qd.LIB_PATH.encode(), path.encode(), 3.0, frames.ctypes.data_as(C.c_void_p), B, I, iters * threads, threads,
                                   first.ctypes.data_as(C.c_void_p), first.size)
    print(f"{threads:2d} threads: {B * iters * threads / secs / 1e3:8.1f} k frames/s", flush=True)
