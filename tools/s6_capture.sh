mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/s6_tests.log 2>&1; tail -3 gpurun_out/s6_tests.log
for b in 512 16384; do
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:input_tc_kernel -c 2 --csv --log-file gpurun_out/s6_tc_$b.csv python tools/profile_step.py --batch $b --steps 2 --warmup 0 > /dev/null 2>&1
echo "batch $b: $(grep -o 'input_tc_kernel.*' gpurun_out/s6_tc_$b.csv | sed 's/(.*gpu__time_duration.sum//' | tr '\n' ' ')"
done
