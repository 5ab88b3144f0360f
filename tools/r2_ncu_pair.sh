#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qlayer_pair -s 7 -c 1 -o gpurun_out/r2i_hidden_stream -f \
    python tools/profile_step.py --batch 16384 --steps 1 --warmup 1 > gpurun_out/r2i_hidden_stream.log 2>&1
tail -3 gpurun_out/r2i_hidden_stream.log; ls -la gpurun_out/r2i_hidden_stream.ncu-rep
