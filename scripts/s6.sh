#!/bin/bash
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms_avg"],4), round(d["roofline"]["frac"],3), d["config"].get("strategy","")[-40:])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
B="python bench.py --no-cpu --no-e2e --steps 20 --warmup 3"
for q in q1 q6; do for d in 0 1 2 3; do
QGPU_FUSED_DBG=$d timeout 120 $B --query $q > $O/s6_${q}_d$d.json 2> $O/s6_${q}_d$d.err; show $O/s6_${q}_d$d.json
done; done
