#!/bin/bash
# usage (under gpurun --gpus N): gpu_r2_n.sh N "<bench args A>" "<bench args B>" ...
set -u
N=$1; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/gpus_n$N.txt 2>&1
nvidia-smi topo -m > $OUT/topo_n$N.txt 2>&1
i=0
for a in "$@"; do
  echo "== N=$N bench $a"
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-ref-work $a > $OUT/r2_n${N}_$i.json 2> $OUT/r2_n${N}_$i.err || tail -20 $OUT/r2_n${N}_$i.err
  grep "gather mode" $OUT/r2_n${N}_$i.err
  python - $OUT/r2_n${N}_$i.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.1f Mrays/s  ms/step %.3f  e2e %.1f (%.3f ms)  launches %d  check: %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["frame_check"]))
for r, row in enumerate(d["per_rank_stage_ms"]):
    print("  rank %d: primary %.2f compact %.2f shade %.2f reflect %.2f resolve %.2f | step %.2f" % (r, *row))
PY
  i=$((i+1))
done
