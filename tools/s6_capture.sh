mkdir -p gpurun_out
for b in 512 16384; do
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:input_fixup_block -s 1 -c 1 -o gpurun_out/s6_fixup_$b -f python tools/profile_step.py --batch $b --steps 1 --warmup 1 > gpurun_out/s6_fixup_$b.log 2>&1; tail -2 gpurun_out/s6_fixup_$b.log
done
