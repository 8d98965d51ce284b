#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for f in 1 2 3; do
  timeout -k 10 400 python bench.py --no-cpu-baseline --no-ref-work --frames-in-flight $f --steps 10 > $OUT/bench_fif$f.json 2> $OUT/bench_fif$f.err || tail -20 $OUT/bench_fif$f.err
  python - $OUT/bench_fif$f.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.1f ms %.3f e2e %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
PY
done
timeout -k 10 400 python bench.py --no-cpu-baseline --no-ref-work --workload cfg5_hair1M_4k --steps 10 > $OUT/bench_fif_hair.json 2> $OUT/bench_fif_hair.err || tail -20 $OUT/bench_fif_hair.err
python - $OUT/bench_fif_hair.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("hair value %.1f ms %.3f e2e %.3f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
PY
