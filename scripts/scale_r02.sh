#!/bin/bash
# The driver's scaling run, reproduced: the default bench line at N GPUs (all workloads under `queries`).  bash scripts/scale_r02.sh N
N=$1
O=gpurun_out
if [ "$N" == "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 50 --warmup 5 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
fi
python - <<P
import json
d=json.loads([l for l in open("$O/r02_bench_n$N.json") if l.startswith("{")][-1])
e=d["e2e"]
print("N=$N q1", round(d["ms_per_step"],4), "%.3e"%d["value"], d["parity_check"]["ok"], "e2e", round(e["ms_per_step"],1), "%.3e"%e["value"], {k:round(v.get("ms_per_step",0),1) for k,v in e["paths"].items()})
for k,v in d["queries"].items(): print("   ",k, round(v.get("ms_per_step",0),4), "%.3e"%v.get("value",0), v.get("error"), v.get("parity_check",{}).get("ok"))
P
