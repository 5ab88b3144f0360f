#!/bin/bash
# round 2, first GPU call: parity suite, int8 peak calibration, D2H ceiling, bench line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
nproc >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1
tail -30 gpurun_out/r2a_pytest.log
timeout 120 ./tools/int8_peak 2 > gpurun_out/r2a_int8_peak.json 2> gpurun_out/r2a_int8_peak.err; cat gpurun_out/r2a_int8_peak.json gpurun_out/r2a_int8_peak.err
timeout 120 python tools/d2h_ceiling.py > gpurun_out/r2a_d2h.json 2>&1; cat gpurun_out/r2a_d2h.json
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json | head -c 3000
