#!/bin/bash
python -m pytest tests -m gpu -x -q -k "fused or config2 or stages_bit_exact or calculate or stress or tensor_core_and_dp4a" 2>&1 | tail -2
python tools/fused_times.py L 2>&1 | grep -E "M= +512|M= +1024"
python tools/fused_timeline.py L 512 2>&1 | grep -E "layer [136]"
