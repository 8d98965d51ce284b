#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout -k 10 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_cull2.log 2>&1; rc=$?; tail -4 $OUT/pytest_cull2.log
echo "== shard probe"; timeout -k 10 400 python scripts/gpu_shard_probe.py --mod 8 --tile 32 --sets "" "19=0" --out $OUT/probe_cull.json 2>&1 | grep -v "^\[ours\]" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('%-10s whole %.3f shard max %.3f mean %.3f eff %.3f' % (d['opts'], d['whole_ms'], d['shard_max_ms'], d['shard_mean_ms'], d['kernel_side_efficiency']))
"
timeout -k 10 300 python bench.py --no-cpu-baseline --no-ref-work --steps 8 > $OUT/bench_cull_final.json 2> $OUT/bench_cull_final.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_cull_final.json") if l.startswith("{")][-1])
print("default bench: %.1f Mrays/s %.3f ms e2e %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["roofline"]["stage_ms_per_step"])
PY
exit $rc
