#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
RTB200_TRACE=1 timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_build or device_transform or work_counters" > $OUT/pytest_build.log 2>&1; tail -3 $OUT/pytest_build.log | cut -c1-300
for w in cfg4_sphere10M_4k_16spp cfg5_hair1M_4k; do
for o in "" "--opt 15=0"; do
echo "== $w $o"
RTB200_TRACE=1 timeout -k 10 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-ref-work $o > $OUT/build4.json 2> $OUT/build4.err || tail -5 $OUT/build4.err
grep "device octree\|octree build\|octree:" $OUT/build4.err | tail -8
python - $OUT/build4.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]; w = d["work"]
print("ms/step %.3f  stages %s  records %d  work: primary V %.0fM T %.0fM shadow V %.0fM T %.0fM" % (d["ms_per_step"], {k: round(v, 3) for k, v in r["stage_ms_per_step"].items()}, d["bvh"]["child_records"], w["primary_volume_tests"]/1e6, w["primary_triangle_tests"]/1e6, w["shadow_volume_tests"]/1e6, w["shadow_triangle_tests"]/1e6))
PY
done; done
