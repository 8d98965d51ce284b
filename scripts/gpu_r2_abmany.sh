#!/bin/bash
# same box: the in-tree library against several variant builds under build/
set -u
OUT=gpurun_out; mkdir -p $OUT
run() {
  timeout -k 10 300 python bench.py --no-cpu-baseline --no-ref-work --frames-in-flight 1 --steps 8 $2 > $OUT/abm_$1.json 2> $OUT/abm_$1.err || tail -5 $OUT/abm_$1.err
  python - $OUT/abm_$1.json "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]["stage_ms_per_step"]
print("%-14s %s: %.3f ms  primary %.3f shade %.3f" % (sys.argv[2], d["config"]["workload"][:10], d["ms_per_step"], r["k_primary"], r["k_shade"]))
PY
}
run base ""
for v in "$@"; do run $v "--lib build/lib_$v.so"; done
run base_again ""
run base_hair "--workload cfg5_hair1M_4k"
for v in "$@"; do run ${v}_hair "--lib build/lib_$v.so --workload cfg5_hair1M_4k"; done
