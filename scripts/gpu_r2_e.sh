#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_lanes or hair_scene" > $OUT/pytest_quick.log 2>&1; rc=$?; tail -3 $OUT/pytest_quick.log
echo "== shard probe tile 32"; timeout -k 10 600 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/shard_probe5.json --sets "" "13=1" "11=0" "13=1,11=0" "13=1,9=2,7=-128" "13=1,10=-32" 2> $OUT/shard_probe5.err | cut -c1-400
