# A/B and switch-off timings of the input layer's fix-up kernels (ncu gpu__time_duration, warm L2)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/fb_tests.log 2>&1; tail -3 gpurun_out/fb_tests.log
for dbg in ${DBGS:-0 3}; do for b in 512 16384; do
FDNN_FB_DEBUG=$dbg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:input_fixup -c 2 --csv --log-file gpurun_out/fb_fb_${dbg}_$b.csv python tools/profile_step.py --batch $b --steps 2 --warmup 0 > /dev/null 2>&1
echo "debug $dbg batch $b: $(grep -o 'input_fixup[a-z_]*kernel.*' gpurun_out/fb_fb_${dbg}_$b.csv | sed 's/(.*gpu__time_duration.sum//' | tr '\n' ' ')"
done; done
timeout 200 python tools/stage_times.py > gpurun_out/fb_stage.log 2>&1; cat gpurun_out/fb_stage.log
