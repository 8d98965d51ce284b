#!/bin/bash
# adaptive (drain) round budgets: parity subset, then whole frame + 1/8 shards on cfg4 and hair for several settings
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or hair_scene or light_space or cfg1" > $OUT/pytest_drain.log 2>&1; tail -3 $OUT/pytest_drain.log
SETS=("" "7=256,10=16,11=0" "7=-128,10=-8,11=-128" "7=-512,10=-32,11=-512" "7=-256,10=-32,11=-256" "7=-256,10=-16,11=0")
echo "== cfg4"; timeout -k 10 900 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --out $OUT/drain_cfg4.json --sets "${SETS[@]}" 2> $OUT/drain_cfg4.err | cut -c1-330
echo "== hair"; timeout -k 10 900 python scripts/gpu_shard_probe.py --workload cfg5_hair1M_4k --mod 8 --tile 32 --out $OUT/drain_hair.json --sets "${SETS[@]}" 2> $OUT/drain_hair.err | cut -c1-330
