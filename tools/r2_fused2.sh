#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/fused_timeline.py L 512 > gpurun_out/r2c_timeline.log 2>&1
python tools/fused_timeline.py S 128 >> gpurun_out/r2c_timeline.log 2>&1
cat gpurun_out/r2c_timeline.log
for dbg in 0 2 4 8 16 30; do echo "FDNN_FUSED_DEBUG=$dbg"; FDNN_FUSED_DEBUG=$dbg timeout 200 python tools/fused_times.py L 2>&1 | grep -E "latency +M= +(128|512)"; done > gpurun_out/r2c_debug.log 2>&1
cat gpurun_out/r2c_debug.log
