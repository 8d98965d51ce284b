#!/bin/bash
# GPU tests, then bench sweeps over a library option; usage: gpu_sweep.sh "<bench args A>" "<bench args B>" ...
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -5 $OUT/pytest_gpu.log
[ $rc -ne 0 ] && exit $rc
i=0
for a in "$@"; do
  echo "== bench $a"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $a > $OUT/sweep_$i.json 2> $OUT/sweep_$i.err || tail -5 $OUT/sweep_$i.err
  python - <<PY
import json
d = json.loads(open("$OUT/sweep_$i.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value %.1f Mrays/s  ms/step %.1f  e2e %.1f  stages %s  frac %.3f  work %s  build %.0f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 1) for k, v in r["stage_ms_per_step"].items()}, r["frac"], d.get("work"), d["bvh"]["build_ms"]))
PY
  i=$((i+1))
done
