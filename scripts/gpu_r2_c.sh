#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== quick parity"
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or hair_scene or packet_and_single or frames_vs_oracle or two_lanes" > $OUT/pytest_quick.log 2>&1; rc=$?; tail -5 $OUT/pytest_quick.log
[ $rc -ne 0 ] && { echo "quick parity failed ($rc)"; tail -40 $OUT/pytest_quick.log; exit $rc; }
echo "== shard probe"; timeout -k 10 600 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/shard_probe3.json --sets "" "12=0" "9=0" "9=3" "7=-128,11=-128" "10=-32" "10=-16" "7=-64,11=-64,10=-16" 2> $OUT/shard_probe3.err | cut -c1-500; tail -3 $OUT/shard_probe3.err
echo "== trace"; timeout -k 10 300 python scripts/gpu_shard_trace.py 8 32 2> $OUT/shard_trace3.txt; grep -n "====\|fused\|split" $OUT/shard_trace3.txt | head -60
