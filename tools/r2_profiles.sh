#!/bin/bash
# round-2 ncu evidence: launch lists of one step (both tile policies) and full captures of the dominant kernels
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
# launch lists (device time per launch; cold-cache and serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 16 --csv --log-file gpurun_out/${TAG}_launches_latency.csv \
    python tools/profile_step.py --steps 4 --warmup 2 > gpurun_out/${TAG}_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 22 -c 22 --csv --log-file gpurun_out/${TAG}_launches_throughput_warm.csv \
    python tools/profile_step.py --steps 2 --warmup 2 --policy throughput >> gpurun_out/${TAG}_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 22 -c 22 --csv --log-file gpurun_out/${TAG}_launches_throughput.csv \
    python tools/profile_step.py --steps 2 --warmup 2 --policy throughput >> gpurun_out/${TAG}_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 11 -c 11 --csv --log-file gpurun_out/${TAG}_launches_stream.csv \
    python tools/profile_step.py --batch 16384 --steps 2 --warmup 1 >> gpurun_out/${TAG}_launches.log 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:qlayer_fused -s 2 -c 1 -o gpurun_out/${TAG}_fused -f \
    python tools/profile_step.py --steps 2 --warmup 2 > gpurun_out/${TAG}_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:qlayer_tc_kernel -s 14 -c 2 -o gpurun_out/${TAG}_hidden_throughput -f \
    python tools/profile_step.py --steps 2 --warmup 2 --policy throughput > gpurun_out/${TAG}_hidden_throughput.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlayer_pair -s 7 -c 1 -o gpurun_out/${TAG}_hidden_stream -f \
    python tools/profile_step.py --batch 16384 --steps 1 --warmup 1 > gpurun_out/${TAG}_hidden_stream.log 2>&1
ls -la gpurun_out/${TAG}_* | tail -12
