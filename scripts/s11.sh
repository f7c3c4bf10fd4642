#!/bin/bash
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print(sys.argv[1], round(d["ms_per_step"],4), round(r["kernel_ms_avg"],4), round(r["frac"],3), round(r["kernel_share_of_step"],3), d["gpu_launches_per_step"], [k["name"] for k in d["kernels"]])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
B="python bench.py --no-cpu --no-e2e --steps 40 --warmup 5"
for q in q1 q6 q3; do
timeout 120 $B --query $q > $O/s11_$q.json 2> $O/s11_$q.err; show $O/s11_$q.json
QGPU_BENCH_PROFILE_ALL=1 timeout 120 $B --query $q > $O/s11_${q}_all.json 2> $O/s11_${q}_all.err; show $O/s11_${q}_all.json
done
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
