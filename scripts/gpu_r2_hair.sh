#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for o in "" "--opt 6=0" "--opt 7=-64 --opt 10=-8" "--opt 7=-32 --opt 11=-64"; do
echo "== hair $o"
RTB200_TRACE=1 timeout -k 10 600 python bench.py --workload cfg5_hair1M_4k --steps 3 --warmup 3 --no-cpu-baseline --no-ref-work $o > $OUT/hair.json 2> $OUT/hair.err
grep "packets\|split" $OUT/hair.err | tail -4 | cut -c1-250
python - $OUT/hair.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]
print("value %.1f Mrays/s traced %.1f  ms/step %.2f  e2e %.1f  stages %s  rays %d traced %d hits %d" % (d["value"], d["traced_mrays_s"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["stage_ms_per_step"].items()}, d["rays_per_step"], d["traced_rays_per_step"], d["work"]["primary_hits"]))
PY
done
