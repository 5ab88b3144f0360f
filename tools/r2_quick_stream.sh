#!/bin/bash
python -m pytest tests -m gpu -x -q -k "pair or stream_chunk" 2>&1 | tail -2
python tools/stream_times.py 16384 2>&1 | tail -2
