#!/bin/bash
# ncu of the two packet kernels of the headline frame (one GPU).  Plain run first, then launch list, then --set full.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2}
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-work --frames-in-flight 1 --opt 18=0"
$BCMD > $OUT/ncu_plain.json 2> $OUT/ncu_plain.err || { tail -20 $OUT/ncu_plain.err; exit 1; }
echo "== ncu launch list"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BCMD > $OUT/ncu_launches.log 2>&1
echo "rc=$?"; tail -2 $OUT/ncu_launches.log | cut -c1-300
echo "== ncu full"
# the packet kernels of ONE timed frame: the instrumented frame (4 launches: 2 chunks x 2 stages) and the 3 warm-up frames are skipped
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:"k_primary_packet|k_shade_packet" -s 16 -c 4 -o $OUT/${TAG}_prof $BCMD > $OUT/ncu_full.log 2>&1
echo "rc=$?"; tail -2 $OUT/ncu_full.log | cut -c1-300
ls -la $OUT | grep $TAG
