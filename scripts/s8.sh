#!/bin/bash
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print(sys.argv[1], round(d["ms_per_step"],4), round(r["kernel_ms_avg"],4), round(r["frac"],3), r.get("read_only_reference"), d["config"].get("strategy","")[-40:])
    for k in d["kernels"][:6]: print("    ", k["name"], round(k["total_ms"]/k["launches"],4), k["launches"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -3
B="python bench.py --no-cpu --no-e2e --steps 30 --warmup 5"
for q in q1 q6 q3; do timeout 120 $B --query $q > $O/s8_$q.json 2> $O/s8_$q.err; show $O/s8_$q.json; done
