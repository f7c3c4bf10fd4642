#!/bin/bash
O=gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print(sys.argv[1], round(d["ms_per_step"],4), round(r["kernel_ms_avg"],4), round(r["frac"],3), d["config"].get("strategy","")[:90], d["step_ms"][:6])
    for k in d["kernels"][:4]: print("    ", k["name"], round(k["total_ms"]/k["launches"],4), k["launches"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_repartition.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-cpu --no-e2e --steps 30 --warmup 5"
timeout 120 $B --query q3 > $O/s9_q3.json 2> $O/s9_q3.err; show $O/s9_q3.json
QGPU_BENCH_SAMPLE_MS=1000 timeout 120 $B --query q3 > $O/s9_q3_nosample.json 2> $O/s9_q3_nosample.err; show $O/s9_q3_nosample.json
QGPU_FUSED_CTAS=3 timeout 120 $B --query q3 > $O/s9_q3_c3.json 2> $O/s9_q3_c3.err; show $O/s9_q3_c3.json
timeout 120 $B --query q3 > $O/s9_q3_b.json 2> $O/s9_q3_b.err; show $O/s9_q3_b.json
