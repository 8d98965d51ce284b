#!/bin/bash
# round-end sequence on one GPU at the final kernels: pytest -m gpu, smoke, default bench (+ cpu baseline), cfg3 bench
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_final.log 2>&1; rc=$?; tail -5 $OUT/pytest_final.log
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench default"; timeout -k 10 900 python bench.py > $OUT/final_n1.json 2> $OUT/final_n1.err || tail -5 $OUT/final_n1.err
echo "== bench reference arm"; timeout -k 10 900 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/final_ref.json 2> $OUT/final_ref.err || tail -5 $OUT/final_ref.err
echo "== bench cfg3"; timeout -k 10 600 python bench.py --workload cfg3_robot_reflect16_1080p --cpu-row-step 2 > $OUT/final_cfg3.json 2> $OUT/final_cfg3.err || tail -5 $OUT/final_cfg3.err
python - <<'PY'
import json
for f in ("final_n1", "final_cfg3", "final_ref"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, d.get("impl"), "%.1f %s  %.3f ms" % (d["value"], d["unit"], d.get("ms_per_step", 0)), "e2e", d["e2e"].get("ms_per_step"), (d.get("roofline") or {}).get("stage_ms_per_step"), (d.get("roofline_issue") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
PY
exit $rc
