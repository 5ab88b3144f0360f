#!/bin/bash
for pdl in 1 0; do echo "FDNN_PDL=$pdl"; FDNN_PDL=$pdl SWEEP_GRIDS=148 SWEEP_LANES=1,2,4,6 python tools/fused_sweep.py 2>&1 | grep -E "layer-by-layer"; done
