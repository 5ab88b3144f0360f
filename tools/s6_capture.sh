mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/s6_tests.log 2>&1; tail -3 gpurun_out/s6_tests.log
timeout 200 python tools/stage_times.py > gpurun_out/s6_stage.log 2>&1; cat gpurun_out/s6_stage.log
POLICY=throughput SWEEP=1,4 timeout 300 python tools/inflight_sweep.py 2>&1 | tail -2
