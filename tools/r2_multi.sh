#!/bin/bash
# N GPUs: device-group tests, torchrun bench, single-process (one handle) bench, D2H ceiling
set -u
N=${1:-2}
TAG=${2:-r2g}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1; nproc >> gpurun_out/${TAG}_topo.txt; numactl -H >> gpurun_out/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "device_group or two_ranks" > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_ceiling.py > gpurun_out/${TAG}_d2h_${N}gpu.json 2> gpurun_out/${TAG}_d2h.err; tail -1 gpurun_out/${TAG}_d2h_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --no-peak > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "rc=$?"; tail -c 500 gpurun_out/${TAG}_bench_${N}gpu.err
timeout 600 python bench.py --single-process --gpus $N --steps 100 > gpurun_out/${TAG}_bench_group_${N}gpu.json 2> gpurun_out/${TAG}_bench_group.err; echo "rc=$?"; tail -c 500 gpurun_out/${TAG}_bench_group.err
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench_${N}gpu.json','gpurun_out/${TAG}_bench_group_${N}gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', d['value'], 'n', d['n_gpus'])
    print(' e2e', {k:v for k,v in d['e2e'].items() if k not in ('mode',)})
    print(' e2e_jni', d.get('e2e_jni'))
    s=d.get('stream1m'); print(' stream1m', s and (s['device_resident']['value'], s['e2e']['value']))
PY
