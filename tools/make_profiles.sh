#!/bin/bash
# Run under gpurun: bench line + ncu launch list + full captures of the two dominant kernels.
# Outputs land in gpurun_out/ (summarised into profiles/ by tools/summarize_profiles.py).
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 400 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# every launch of two steps (after one warm-up step = 11 launches: 3 input + 7 int8 layers + softmax), device time only
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 11 -c 22 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/profile_step.py --steps 2 --warmup 1 > gpurun_out/${TAG}_launches.log 2>&1
# same with warm caches (ncu's default flushes L2 between kernels, which is not how the pass runs)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 11 -c 22 --csv \
    --log-file gpurun_out/${TAG}_launches_warm.csv python tools/profile_step.py --steps 2 --warmup 1 >> gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlayer_tc -s 7 -c 2 -o gpurun_out/${TAG}_hidden -f \
    python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/${TAG}_hidden.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:input_tc_kernel -s 1 -c 1 -o gpurun_out/${TAG}_input -f \
    python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/${TAG}_input.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:input_fixup_block -s 1 -c 1 -o gpurun_out/${TAG}_fixup -f \
    python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/${TAG}_fixup.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:input_fixup_block -s 1 -c 1 -o gpurun_out/${TAG}_fixup_stream -f \
    python tools/profile_step.py --batch 16384 --steps 1 --warmup 1 > gpurun_out/${TAG}_fixup_stream.log 2>&1
# long-stream regime (16384-frame chunks): every int8 layer runs on CTA pairs (qlayer_pair.cu); first hidden layer and the output layer
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlayer_pair -s 7 -c 1 -o gpurun_out/${TAG}_hidden_stream -f \
    python tools/profile_step.py --batch 16384 --steps 1 --warmup 1 > gpurun_out/${TAG}_hidden_stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qlayer_pair -s 13 -c 1 -o gpurun_out/${TAG}_output_stream -f \
    python tools/profile_step.py --batch 16384 --steps 1 --warmup 1 > gpurun_out/${TAG}_output_stream.log 2>&1
timeout 300 python tools/stage_times.py > gpurun_out/${TAG}_stage_times.log 2>&1
ls -la gpurun_out | tail -12
