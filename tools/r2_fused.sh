#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or config2 or config3 or headline or calculate or shared_model or variable_length" > gpurun_out/r2b_pytest.log 2>&1
tail -15 gpurun_out/r2b_pytest.log
timeout 300 python tools/fused_times.py L > gpurun_out/r2b_fused_times.log 2>&1; cat gpurun_out/r2b_fused_times.log
timeout 300 python tools/fused_times.py S >> gpurun_out/r2b_fused_times.log 2>&1; tail -6 gpurun_out/r2b_fused_times.log
