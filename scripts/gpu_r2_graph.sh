#!/bin/bash
# frame graphs: parity tests, then one rank's share of the 8-way sharded frame with and without them
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_graph.log 2>&1; rc=$?; tail -15 $OUT/pytest_graph.log
echo "== shard probe"; timeout -k 10 400 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --sets "" "18=0" --out $OUT/probe_graph.json 2>&1 | grep -v "^\[ours\]" | tail -4
echo "== pipeline probe"; timeout -k 10 400 python scripts/gpu_pipeline_probe.py --contexts 2 --out $OUT/pipeline_graph.json 2>&1 | grep -v "^\[ours\]" | tail -6
exit $rc
