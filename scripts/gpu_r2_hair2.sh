#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for o in "$@"; do
echo "== hair $o"
RTB200_TRACE=1 timeout -k 10 600 python bench.py --workload cfg5_hair1M_4k --steps 3 --warmup 3 --no-cpu-baseline --no-ref-work $o > $OUT/hair.json 2> $OUT/hair.err || tail -5 $OUT/hair.err
grep "shadow packets\|split shadow" $OUT/hair.err | tail -2 | cut -c1-250
python - $OUT/hair.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]; w = d["work"]; h = max(w["primary_hits"], 1)
print("value %.1f Mrays/s  ms/step %.2f  stages %s | per shadow ray: V %.0f T %.0f" % (d["value"], d["ms_per_step"], {k: round(v, 2) for k, v in r["stage_ms_per_step"].items()}, w["shadow_volume_tests"] / h, w["shadow_triangle_tests"] / h))
PY
done
