#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  for fused in 2 0; do
    echo "== $tool FDNN_FUSED=$fused"
    FDNN_FUSED=$fused FDNN_GRAPHS=0 timeout 600 compute-sanitizer --tool $tool $( [ $tool = synccheck ] && echo --num-cuda-barriers 65536 ) python tools/sanitizer_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rows sum|lazy row|Error|hazard" | head -8
  done
done
