#!/bin/bash
# ncu --set full of the heavy (sphere) chunk's k_primary and k_shade launches of one bench frame (separate captures).
set -u
OUT=gpurun_out; mkdir -p $OUT
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${2:-}"
timeout 600 $BCMD > $OUT/plain3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_primary" -s ${1:-5} -c 1 -o $OUT/prof_primary -f $BCMD > $OUT/ncu_primary.log 2>&1
echo "rc=$?"; tail -2 $OUT/ncu_primary.log
timeout 600 $BCMD > $OUT/plain4.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s ${1:-5} -c 1 -o $OUT/prof_shade -f $BCMD > $OUT/ncu_shade.log 2>&1
echo "rc=$?"; tail -2 $OUT/ncu_shade.log; ls -la $OUT/prof_*
