bash tools/make_profiles.sh r1f > gpurun_out/r1f_make.log 2>&1; tail -3 gpurun_out/r1f_make.log
echo "== synccheck"; timeout 280 compute-sanitizer --tool synccheck --num-cuda-barriers 8192 --print-limit 6 python tools/sanitizer_case.py 2>&1 | grep -v "^$" | grep -v "Host Frame\|Saved host" | tail -8
