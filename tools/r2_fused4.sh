#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or config2 or config3 or headline or calculate or shared_model or variable_length or lazy" > gpurun_out/r2e_pytest.log 2>&1
tail -5 gpurun_out/r2e_pytest.log
timeout 300 python tools/fused_times.py L > gpurun_out/r2e_fused_times.log 2>&1; cat gpurun_out/r2e_fused_times.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-stream > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 300 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'single',d['single_stream'],'e2e',d['e2e']['value'],d['e2e']['single_caller_ms_per_step'],'launches',d['gpu_launches'])
PY
