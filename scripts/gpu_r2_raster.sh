#!/bin/bash
# raster_trace on the GPU + the e2e timing of the default bench
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest raster"; timeout -k 10 900 python -m pytest tests -m gpu -x -q -k "raster or adapter or call_order or unsupported" > $OUT/pytest_raster.log 2>&1; rc=$?; tail -25 $OUT/pytest_raster.log
echo "== bench default (no cpu baseline)"; timeout -k 10 600 python bench.py --no-cpu-baseline --no-ref-work > $OUT/bench_e2e.json 2> $OUT/bench_e2e.err; tail -4 $OUT/bench_e2e.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_e2e.json") if l.startswith("{")][-1])
print("value %.1f ms %.3f e2e %s" % (d["value"], d["ms_per_step"], d["e2e"]))
PY
exit $rc
