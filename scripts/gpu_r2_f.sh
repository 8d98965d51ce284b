#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -6 $OUT/pytest_gpu.log
[ $rc -ne 0 ] && exit $rc
echo "== shard probe"; timeout -k 10 600 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/shard_probe6.json --sets "" 2> $OUT/shard_probe6.err | cut -c1-400
echo "== trace"; timeout -k 10 300 python scripts/gpu_shard_trace.py 8 32 2> $OUT/shard_trace6.txt; grep -n "====" -B12 $OUT/shard_trace6.txt | grep "whole" -B14 | grep -v "items" | head -30
