#!/bin/bash
# One gpurun call: GPU parity tests, packet diagnostics, kernel A/B sweep, full bench, ncu launch list + full captures.
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -5 $OUT/pytest_gpu.log
[ $rc -ne 0 ] && { echo "gpu tests failed ($rc)"; exit $rc; }
summ() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]
print("value %.1f Mrays/s  ms/step %.2f  e2e %.1f  stages %s  frac %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["stage_ms_per_step"].items()}, r["frac"], d["gpu_launches"]))
PY
}
i=0
for a in "$@"; do
  echo "== bench $a"
  RTB200_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-work $a > $OUT/ab_$i.json 2> $OUT/ab_$i.err || tail -5 $OUT/ab_$i.err
  grep "packets" $OUT/ab_$i.err | head -2
  summ $OUT/ab_$i.json
  i=$((i+1))
done
[ "${FULL:-0}" = "1" ] || exit 0
echo "== full bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err || { tail -20 $OUT/bench.err; exit 1; }; summ $OUT/bench.json; tail -3 $OUT/bench.err
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-work"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $BCMD > $OUT/ncu_launches.log 2>&1
echo "rc=$?"; tail -2 $OUT/ncu_launches.log | cut -c1-300
echo "== ncu full"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_primary_packet|k_shade_packet" -s 8 -c 2 -o $OUT/prof_r1 $BCMD > $OUT/ncu_full.log 2>&1
echo "rc=$?"; tail -2 $OUT/ncu_full.log | cut -c1-300
ls -la $OUT | tail -20
