#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest fans"; timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "fan or cfg3 or reflect or mirror" > $OUT/pytest_fan.log 2>&1; rc=$?; tail -15 $OUT/pytest_fan.log
for o in 1 0; do
timeout -k 10 300 python bench.py --workload cfg3_robot_reflect16_1080p --no-cpu-baseline --no-ref-work --frames-in-flight 1 --opt 20=$o --steps 8 > $OUT/fan_$o.json 2> $OUT/fan_$o.err || tail -5 $OUT/fan_$o.err
python - $OUT/fan_$o.json $o <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("fan lanes %s: %.3f ms  %.1f Mrays/s  stages %s" % (sys.argv[2], d["ms_per_step"], d["value"], {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()}))
PY
done
exit $rc
