#!/bin/bash
# N-GPU confirmation run (gpurun --gpus N): the three sharded benches, short
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
for q in q1 q3 groupby; do
  extra=""; [ $q != q1 ] && extra="--no-e2e --no-cpu"
  [ $q == q1 ] && extra="--e2e-steps 2"
  timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --query $q $extra > gpurun_out/bench_${q}_n$N.json 2> gpurun_out/bench_${q}_n$N.err
  tail -1 gpurun_out/bench_${q}_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$q', d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['config']['strategy'][:70], [(k['name'],round(k['total_ms']/5,2)) for k in d['kernels'][:5]])
except Exception as e: print('$q FAILED', e)"
  grep -i "error\|Traceback" gpurun_out/bench_${q}_n$N.err | head -3
done
