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
python scripts/readbw.py
B="python bench.py --no-cpu --no-e2e --steps 20 --warmup 3"
QGPU_FUSED_DBG=0 timeout 120 $B --query q6 > $O/s7_q6_base.json 2> $O/s7_q6_base.err; show $O/s7_q6_base.json
QGPU_FUSED_CTAS=2 timeout 120 $B --query q6 > $O/s7_q6_c2.json 2> $O/s7_q6_c2.err; show $O/s7_q6_c2.json
QGPU_FUSED_CTAS=2 QGPU_FUSED_DBG=1 timeout 120 $B --query q6 > $O/s7_q6_c2d1.json 2> $O/s7_q6_c2d1.err; show $O/s7_q6_c2d1.json
QGPU_FUSED_DBG=4 timeout 120 $B --query q6 > $O/s7_q6_d4.json 2> $O/s7_q6_d4.err; show $O/s7_q6_d4.json
QGPU_FUSED_DBG=5 timeout 120 $B --query q6 > $O/s7_q6_d5.json 2> $O/s7_q6_d5.err; show $O/s7_q6_d5.json
QGPU_FUSED_DBG=0 timeout 120 $B --query q1 > $O/s7_q1_base.json 2> $O/s7_q1_base.err; show $O/s7_q1_base.json
QGPU_FUSED_DBG=4 timeout 120 $B --query q1 > $O/s7_q1_d4.json 2> $O/s7_q1_d4.err; show $O/s7_q1_d4.json
QGPU_FUSED_DBG=5 timeout 120 $B --query q1 > $O/s7_q1_d5.json 2> $O/s7_q1_d5.err; show $O/s7_q1_d5.json
