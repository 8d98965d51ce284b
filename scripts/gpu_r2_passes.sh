#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
SETS=("$@")
echo "== cfg4"; timeout -k 10 900 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/passes_cfg4.json --sets "${SETS[@]}" 2> $OUT/passes_cfg4.err | cut -c1-330
echo "== hair"; timeout -k 10 900 python scripts/gpu_shard_probe.py --workload cfg5_hair1M_4k --mod 8 --tile 32 --out $OUT/passes_hair.json --sets "${SETS[@]}" 2> $OUT/passes_hair.err | cut -c1-330
