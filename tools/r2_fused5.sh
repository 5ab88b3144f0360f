#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or config2 or config3 or headline or calculate or shared_model or variable_length or lazy" > gpurun_out/r2k_pytest.log 2>&1; tail -4 gpurun_out/r2k_pytest.log
timeout 300 python tools/fused_times.py L 2>&1 | grep latency | tee gpurun_out/r2k_fused_times.log
SWEEP_GRIDS=64,74 SWEEP_LANES=1,2,4 timeout 600 python tools/fused_sweep.py 2>&1 | tee gpurun_out/r2k_fused_sweep.log
