#!/bin/bash
# same box, two builds: build/librtb200_base.so (HEAD) against the working tree's library; usage: gpu_r2_ablib.sh "<opts for new>" ...
set -u
OUT=gpurun_out; mkdir -p $OUT
run() {  # label, extra args
  timeout -k 10 300 python bench.py --no-cpu-baseline --no-ref-work --frames-in-flight 1 --steps 8 $2 > $OUT/ab_$1.json 2> $OUT/ab_$1.err || tail -5 $OUT/ab_$1.err
  python - $OUT/ab_$1.json "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]["stage_ms_per_step"]
print("%-22s %s: %.3f ms  primary %.3f shade %.3f" % (sys.argv[2], d["config"]["workload"][:10], d["ms_per_step"], r["k_primary"], r["k_shade"]))
PY
}
for rep in 1 2; do
run base_$rep "--lib build/librtb200_base.so"
run new_cull3_$rep ""
run new_cull0_$rep "--opt 19=0"
done
run base_hair "--lib build/librtb200_base.so --workload cfg5_hair1M_4k"
run new_hair "--workload cfg5_hair1M_4k"
