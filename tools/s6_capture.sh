mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/s6_tests.log 2>&1; tail -3 gpurun_out/s6_tests.log
timeout 200 python tools/stage_times.py > gpurun_out/s6_stage.log 2>&1; cat gpurun_out/s6_stage.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s6_bench.json')); print(d['value'], d['ms_per_step'], d['single_stream'], d['e2e']['value'], d['stream_regime']['frames_per_s'])"
