#!/bin/bash
# BASELINE.json configs[4]: TPC-H SF100 Q1/Q6/Q3 with lineitem row-range sharded across 8 B200 (gpurun --gpus 8)
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for q in q1 q6 q3; do
  extra="--no-cpu --e2e-steps 2"; [ $q == q3 ] && extra="--no-cpu --no-e2e"
  timeout 600 $TR bench.py --gpus $N --sf 12.5 --steps 10 --warmup 3 --query $q $extra > gpurun_out/r01_bench_${q}_n8_sf100.json 2> gpurun_out/r01_bench_${q}_n8_sf100.err
  tail -1 gpurun_out/r01_bench_${q}_n8_sf100.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$q', d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['config']['workload'][:90], d['roofline']['frac'])
except Exception as e: print('$q FAILED', e)"
  grep -i "error\|Traceback" gpurun_out/r01_bench_${q}_n8_sf100.err | head -3
done
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --query groupby --no-e2e --no-cpu > gpurun_out/r01_bench_groupby_n8.json 2> gpurun_out/r01_bench_groupby_n8.err
tail -1 gpurun_out/r01_bench_groupby_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('groupby', d['value'], d['ms_per_step'], d['step_ms'])"
