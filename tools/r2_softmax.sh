#!/bin/bash
set -u
python -m pytest tests -m gpu -x -q -k "config2 or config3 or calculate or fused or lazy or softmax" 2>&1 | tail -3
python tools/fused_times.py L 2>&1 | grep -E "latency +M= +512"
python tools/stream_times.py 16384 2>&1 | tail -2
python - <<'PY'
import sys; sys.path.insert(0,'/root/repo' if False else '.')
import torch, numpy as np
import fast_dnn_b200
from fast_dnn_b200 import quantized_dnn as qd, synth
dnn=qd.QuantizedDnn.load_from_file(synth.network_file("L"),device=0); dnn.set_tile_policy("throughput")
x=torch.from_numpy(synth.make_frames(512,440,seed=1)).cuda(); y=torch.empty(512,8000,device="cuda")
c=dnn.get_new_lazy_context(512); ms=c.profile_stages(x.data_ptr(),512,y.data_ptr(),iters=50); ms=c.profile_stages(x.data_ptr(),512,y.data_ptr(),iters=100)
print("throughput-policy stages us:", [round(float(v)*1e3,1) for v in ms])
PY
