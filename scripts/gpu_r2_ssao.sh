#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ssao or call_order or cfg1 or analytic" > $OUT/pytest_ssao.log 2>&1; tail -30 $OUT/pytest_ssao.log | cut -c1-250
