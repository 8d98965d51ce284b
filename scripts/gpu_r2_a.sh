#!/bin/bash
# Round-2 first GPU call: parity tests (incl. the BASELINE-size ones), the default bench line, shard probe.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -15 $OUT/pytest_gpu.log
echo "== bench default"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -5 $OUT/bench.err; cut -c1-1500 $OUT/bench.json
echo "== shard probe"; timeout 900 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/shard_probe.json --sets "" "9=3" "9=4" "7=64,10=16,11=64" "7=32,10=16,11=32" "9=4,7=64,10=16,11=64" 2> $OUT/shard_probe.err | cut -c1-600; tail -3 $OUT/shard_probe.err
exit $rc
