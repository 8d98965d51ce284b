#!/bin/bash
# same box: the in-tree library against build/<name>.so given as $1, on the headline frame and the hair frame
set -u
OUT=gpurun_out; mkdir -p $OUT
VAR=$1
run() {
  timeout -k 10 300 python bench.py --no-cpu-baseline --no-ref-work --frames-in-flight 1 --steps 8 $2 > $OUT/ab2_$1.json 2> $OUT/ab2_$1.err || tail -5 $OUT/ab2_$1.err
  python - $OUT/ab2_$1.json "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]["stage_ms_per_step"]
print("%-14s %s: %.3f ms  primary %.3f shade %.3f  tri tests %d + %d" % (sys.argv[2], d["config"]["workload"][:10], d["ms_per_step"], r["k_primary"], r["k_shade"], d["work"]["primary_triangle_tests"], d["work"]["shadow_triangle_tests"]))
PY
}
for rep in 1 2; do
run base_$rep ""
run var_$rep "--lib $VAR"
done
run base_hair "--workload cfg5_hair1M_4k"
run var_hair "--lib $VAR --workload cfg5_hair1M_4k"
run base_fill "--workload cfg4_fill_sphere10M_4k_16spp"
run var_fill "--lib $VAR --workload cfg4_fill_sphere10M_4k_16spp"
