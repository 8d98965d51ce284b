#!/bin/bash
# light-space queue order: parity subset, then hair and cfg4 with the sort on / off
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "light_space or hair or two_lanes" > $OUT/pytest_sort.log 2>&1; tail -5 $OUT/pytest_sort.log
bash scripts/gpu_r2_hair2.sh "" "--opt 14=0" "--opt 7=-64" "--opt 7=-1024 --opt 10=-64"
for o in "" "--opt 14=0"; do
echo "== cfg4 $o"
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-work $o > $OUT/sort4.json 2> $OUT/sort4.err || tail -5 $OUT/sort4.err
python - $OUT/sort4.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]
print("value %.1f Mrays/s  ms/step %.3f  e2e %.3f ms  stages %s launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in r["stage_ms_per_step"].items()}, d["gpu_launches"]))
PY
done
