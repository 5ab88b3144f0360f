#!/usr/bin/env python
"""File-to-file front end (fdnn_calculate_file, csrc/stream_file.cc) on the headline network: frames/s from a big-endian
feature file to (a) /dev/null — reader, byte swap, PCIe and GPU only — and (b) a real binary dump on local storage.

  python tools/file_stream_bench.py [frames] [out_dir] [nodump]     → one JSON line   (FDNN_FILE_DEBUG=1: stage times on stderr)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
out_dir = sys.argv[2] if len(sys.argv) > 2 else "/tmp"
feats = os.path.join(out_dir, "fdnn_stream_features.bin")
dump = os.path.join(out_dir, "fdnn_stream_scores.bin")

block = synth.make_frames(8192, 440, seed=5)
with open(feats, "wb") as f:
    f.write(np.array([frames, 440], dtype=">i4").tobytes())
    be = block.astype(">f4").tobytes()
    for i in range(0, frames, 8192):
        f.write(be[: min(8192, frames - i) * 440 * 4])

dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=0)
res = {"tool": "file_stream_bench", "network": "440-7x2048-8000", "frames": frames, "feature_file_mb": round(os.path.getsize(feats) / 1e6, 1)}
dnn.calculate_file(feats, "/dev/null", chunk_frames=2048)  # warm-up: workspaces, graphs, page cache of the feature file
legs = [("dev_null", "/dev/null", 2048), ("dev_null_chunk4096", "/dev/null", 4096), ("dev_null_chunk512", "/dev/null", 512)]
if "nodump" not in sys.argv[3:]:
    legs.append(("local_file", dump, 2048))
for name, target, chunk in legs:
    t0 = time.perf_counter()
    n = dnn.calculate_file(feats, target, chunk_frames=chunk)
    dt = time.perf_counter() - t0
    assert n == frames
    res[name] = {"frames_per_s": round(frames / dt), "seconds": round(dt, 3), "chunk_frames": chunk}
    if target == dump:
        res[name]["dump_gb"] = round(os.path.getsize(dump) / 1e9, 2)
        with open(dump, "rb") as f:
            assert tuple(np.frombuffer(f.read(8), dtype=np.uint32)) == (frames, 8000)
            got = np.fromfile(f, dtype=np.float32, count=min(frames, 8192) * 8000).reshape(-1, 8000)
        want = dnn.calculate(block[: got.shape[0]])
        res[name]["first_8192_rows_equal_calculate"] = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
        os.remove(dump)
os.remove(feats)
dnn.delete()
print(json.dumps(res))
