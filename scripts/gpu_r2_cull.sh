#!/bin/bash
# bounding-pyramid cull: parity tests, then A/B on the headline frame and the hair frame (RT_OPT_PACKET_CULL 0 / 1 / 2 / 3)
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_cull.log 2>&1; rc=$?; tail -8 $OUT/pytest_cull.log
for w in cfg4_sphere10M_4k_16spp cfg5_hair1M_4k; do
for c in 0 1 2 3; do
  timeout -k 10 300 python bench.py --workload $w --no-cpu-baseline --no-ref-work --frames-in-flight 1 --opt 19=$c --steps 8 > $OUT/cull_${w}_$c.json 2> $OUT/cull_${w}_$c.err || tail -5 $OUT/cull_${w}_$c.err
  python - $OUT/cull_${w}_$c.json $c <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]["stage_ms_per_step"]
print("%s cull=%s: %.3f ms  primary %.3f shade %.3f  vol tests %d + %d" % (d["config"]["workload"], sys.argv[2], d["ms_per_step"], r["k_primary"], r["k_shade"], d["work"]["primary_volume_tests"], d["work"]["shadow_volume_tests"]))
PY
done
done
exit $rc
