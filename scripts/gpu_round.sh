#!/bin/bash
# One gpurun call: GPU parity tests -> smoke -> bench -> ncu launch list -> ncu full capture of the top kernels.
# Every ncu step runs only after the same command exited 0 without ncu (B200_PROFILING.md).
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt

echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -15 $OUT/pytest_gpu.log
[ $rc -ne 0 ] && { echo "gpu tests failed ($rc): skipping profiling"; exit $rc; }
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1 || { tail -20 $OUT/smoke.log; exit 1; }; tail -4 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err || { tail -20 $OUT/bench.err; exit 1; }; cat $OUT/bench.json; tail -5 $OUT/bench.err
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
echo "== ncu launch list"
timeout 600 $BCMD > $OUT/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $BCMD > $OUT/ncu_launches.log 2>&1
echo "rc=$?"; tail -3 $OUT/ncu_launches.log
echo "== ncu full"
timeout 600 $BCMD > $OUT/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_primary|k_shade" -s 14 -c 2 -o $OUT/prof $BCMD > $OUT/ncu_full.log 2>&1
echo "rc=$?"; tail -3 $OUT/ncu_full.log
ls -la $OUT
