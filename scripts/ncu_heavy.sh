#!/bin/bash
# ncu --set full of the heavy (sphere) chunk's k_primary and k_shade launches of one bench frame.
set -u
OUT=gpurun_out; mkdir -p $OUT
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $BCMD > $OUT/plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_primary|k_shade" -s ${1:-10} -c 2 -o $OUT/prof_heavy $BCMD > $OUT/ncu_heavy.log 2>&1
echo "rc=$?"; tail -3 $OUT/ncu_heavy.log; ls -la $OUT/prof_heavy*
