#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
run() {
timeout -k 10 300 python bench.py --workload cfg3_robot_reflect16_1080p --no-cpu-baseline --no-ref-work --frames-in-flight 1 --steps 8 $2 > $OUT/fan2_$1.json 2> $OUT/fan2_$1.err || tail -5 $OUT/fan2_$1.err
python - $OUT/fan2_$1.json $1 <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("%s: %.3f ms  reflect %.3f" % (sys.argv[2], d["ms_per_step"], d["roofline"]["stage_ms_per_step"]["k_reflect"]))
PY
}
run base ""
for v in f4 f5 f6; do run $v "--lib build/lib_$v.so"; done
