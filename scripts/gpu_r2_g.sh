#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -6 $OUT/pytest_gpu.log
[ $rc -ne 0 ] && exit $rc
echo "== shard probe"; timeout -k 10 600 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/shard_probe7.json --sets "" "13=1" 2> $OUT/shard_probe7.err | cut -c1-400
