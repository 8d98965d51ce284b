#!/bin/bash
# the driver's round-end sequence on one GPU: pytest -m gpu, smoke, default bench (+ reference arm), other workloads
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; rc=$?; tail -6 $OUT/pytest_gpu.log
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench default"; timeout -k 10 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err || tail -5 $OUT/bench_n1.err
for w in cfg4_fill_sphere10M_4k_16spp cfg5_hair1M_4k cfg1_robot_720p cfg2_robot_textured_720p_4spp cfg3_robot_reflect16_1080p; do
  echo "== bench $w"; timeout -k 10 900 python bench.py --workload $w --cpu-row-step 2 > $OUT/bench_$w.json 2> $OUT/bench_$w.err || tail -5 $OUT/bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    r = d["roofline"]; c = d.get("cpu_baseline") or {}
    print("%s: %.1f Mrays/s (traced %.1f)  %.3f ms  e2e %.3f ms  stages %s  frac %.3f issue %s cpu %s x%s" % (d["config"]["workload"], d["value"], d["traced_mrays_s"], d["ms_per_step"], d["e2e"]["ms_per_step"],
          {k: round(v, 3) for k, v in r["stage_ms_per_step"].items()}, r["frac"], (d.get("roofline_issue") or {}).get("frac"), c.get("value"), c.get("cores")))
PY
exit $rc
