#!/bin/bash
# fused item scheduling: parity first (short timeout: a scheduling bug would spin), then shard probe fused vs passes
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== quick parity (split packets, fused and passes)"
timeout -k 10 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or hair_scene or packet_and_single or frames_vs_oracle" > $OUT/pytest_quick.log 2>&1; rc=$?; tail -5 $OUT/pytest_quick.log
[ $rc -ne 0 ] && { echo "quick parity failed ($rc)"; tail -40 $OUT/pytest_quick.log; exit $rc; }
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -12 $OUT/pytest_gpu.log
[ $rc -ne 0 ] && exit $rc
echo "== shard probe"; timeout -k 10 600 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/shard_probe2.json --sets "" "12=0" "9=3" "9=4" "7=-128,11=-128" "7=-128,11=-128,10=-32" "7=-512,11=-512" 2> $OUT/shard_probe2.err | cut -c1-500; tail -3 $OUT/shard_probe2.err
echo "== trace"; timeout -k 10 300 python scripts/gpu_shard_trace.py 8 32 2> $OUT/shard_trace2.txt; grep -n "====\|fused\|split" $OUT/shard_trace2.txt | head -40
