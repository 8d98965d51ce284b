#!/bin/bash
# auto light-space order + round-budget sweep on hair and cfg4
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "light_space or hair_scene" > $OUT/pytest_sort.log 2>&1; tail -3 $OUT/pytest_sort.log
bash scripts/gpu_r2_hair2.sh "" "--opt 7=-512 --opt 10=-32" "--opt 7=-1024 --opt 10=-64" "--opt 7=-2048 --opt 10=-64" "--opt 7=-1024 --opt 10=-128" "--opt 7=-1024 --opt 10=-64 --opt 11=1024"
for o in "" "--opt 7=-512 --opt 10=-32" "--opt 7=-1024 --opt 10=-64" "--opt 7=-512 --opt 10=-16"; do
echo "== cfg4 $o"
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-work $o > $OUT/sort4.json 2> $OUT/sort4.err || tail -5 $OUT/sort4.err
python - $OUT/sort4.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d["roofline"]
print("value %.1f Mrays/s  ms/step %.3f  e2e %.3f ms  stages %s launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in r["stage_ms_per_step"].items()}, d["gpu_launches"]))
PY
done
