#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pair or stream_chunk or chunked or config2" > gpurun_out/r2h_pytest.log 2>&1; tail -4 gpurun_out/r2h_pytest.log
for sub in 0 2304 4096 1024; do FDNN_OUTPUT_SUB_ROWS=$sub timeout 200 python tools/stream_times.py 16384; done > gpurun_out/r2h_stream_times.log 2>&1
cat gpurun_out/r2h_stream_times.log
