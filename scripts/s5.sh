#!/bin/bash
# scratch experiment driver (one GPU): packed/compact Q1 variants, sampler period on Q3, ncu capture of the Q1 kernel
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_quirks.py -m gpu -x -q 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms_avg"],4), round(d["roofline"]["frac"],3), d["gpu_launches_per_step"], d["config"].get("strategy","")[-60:], d["step_ms"][:4], d["clocks"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
B="python bench.py --no-cpu --no-e2e --steps 40 --warmup 5"
timeout 120 $B --query q1 > $O/s5_q1.json 2> $O/s5_q1.err; show $O/s5_q1.json
QGPU_FUSED_STAGES=3 timeout 120 $B --query q1 > $O/s5_q1_st3.json 2> $O/s5_q1_st3.err; show $O/s5_q1_st3.json
QGPU_FUSED_NOPACK=1 timeout 120 $B --query q1 > $O/s5_q1_nopack.json 2> $O/s5_q1_nopack.err; show $O/s5_q1_nopack.json
timeout 120 $B --query q6 > $O/s5_q6.json 2> $O/s5_q6.err; show $O/s5_q6.json
timeout 120 $B --query q3 > $O/s5_q3.json 2> $O/s5_q3.err; show $O/s5_q3.json
QGPU_BENCH_SAMPLE_MS=5 timeout 120 $B --query q3 > $O/s5_q3_s5.json 2> $O/s5_q3_s5.err; show $O/s5_q3_s5.json
timeout 300 ncu --clock-control none --set full --import-source on -k regex:k_fused_scan_agg_spec -c 1 -o $O/r01b_q1_fused_spec python bench.py --query q1 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f_q1b.log 2>&1
tail -2 $O/ncu_f_q1b.log
