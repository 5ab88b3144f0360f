#!/bin/bash
# full bench line + reference arm at N = 1
set -u
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"; tail -c 800 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
for k in ('value','ms_per_step','timed','gpu_launches','clocks'): print(k, d[k])
print('single', d['single_stream'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='mode'})
print('e2e_jni', d['e2e_jni'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('kernel','how')})
print('roofline_stream', d['roofline_stream'])
print('lazy', json.dumps(d['lazy'])[:1500])
print('stream1m', json.dumps(d['stream1m'])[:900])
print('configs1', d.get('configs1'))
print('cpu', d['cpu_baseline'])
print('i8', d['int8_peak_calibration'])
PY
cat gpurun_out/${TAG}_bench_reference.json | head -c 700
