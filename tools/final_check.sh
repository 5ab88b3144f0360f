#!/bin/bash
# what the driver runs at round end, in one call: smoke, the GPU suite, both bench arms
set -u
TAG=${1:-r2z}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/${TAG}_pytest.log
bash tools/r2_bench.sh ${TAG}
