#!/bin/bash
# round-end validation in one call: smoke, the whole GPU suite (no -x: every failure is reported), both bench arms,
# the file front end's own bench, and the pair kernel with its saturation scan switched off (what the scan costs now)
set -u
TAG=${1:-r2final}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/${TAG}_pytest.log
bash tools/r2_bench.sh ${TAG}
timeout 300 python tools/file_stream_bench.py 65536 /tmp > gpurun_out/${TAG}_file_stream.json 2> gpurun_out/${TAG}_file_stream.err; echo "file_stream rc=$?"; cat gpurun_out/${TAG}_file_stream.json; tail -3 gpurun_out/${TAG}_file_stream.err
df -h /tmp | tail -1
( timeout 200 python tools/stream_times.py 16384; FDNN_DEBUG=1 timeout 200 python tools/stream_times.py 16384 ) > gpurun_out/${TAG}_scan_off.log 2>&1; cat gpurun_out/${TAG}_scan_off.log
