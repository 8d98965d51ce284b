#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== quick parity"
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or hair_scene or two_lanes" > $OUT/pytest_quick.log 2>&1; rc=$?; tail -3 $OUT/pytest_quick.log
[ $rc -ne 0 ] && { tail -40 $OUT/pytest_quick.log; exit $rc; }
for t in 32 16 64; do
echo "== shard probe tile $t"; timeout -k 10 600 python scripts/gpu_shard_probe.py --mod 8 --tile $t --out $OUT/shard_probe4_$t.json --sets "" "10=-16" "10=-8" "10=-16,9=3" "10=-16,7=-128,11=-128" "10=-16,9=3,7=-128,11=-128" "12=1" 2> $OUT/shard_probe4.err | cut -c1-400
done
