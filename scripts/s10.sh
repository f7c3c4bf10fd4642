#!/bin/bash
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["ms_per_step"],3), [(k["name"], round(k["total_ms"]/k["launches"],3)) for k in d["kernels"][:6]])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_radix.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --query groupby --rows 200000000 --groups 20000000"
for v in 0 8 1 2 4; do
QGPU_RADIX_NOPF=$v timeout 200 $B > $O/s10_gb_$v.json 2> $O/s10_gb_$v.err; show $O/s10_gb_$v.json
done
