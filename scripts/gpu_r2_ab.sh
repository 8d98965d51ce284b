#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg1 or cfg2 or cfg3 or hair_scene or split" > $OUT/pytest_ab.log 2>&1; tail -3 $OUT/pytest_ab.log | cut -c1-250
for w in cfg4_sphere10M_4k_16spp cfg5_hair1M_4k cfg4_fill_sphere10M_4k_16spp; do
echo "== $w"; timeout -k 10 600 python scripts/gpu_ab_libs.py --workload $w "$@" 2> $OUT/ab.err || tail -5 $OUT/ab.err
done
