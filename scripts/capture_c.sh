#!/bin/bash
# N=1 bench lines only (one GPU)
TAG=${1:-r01}
O=gpurun_out
for q in q1 q6 q3; do
  extra=""; [ $q == q3 ] && extra="--no-cpu --steps 30"; [ $q == q6 ] && extra="--steps 30"
  timeout 300 python bench.py --query $q $extra > $O/${TAG}_bench_${q}_n1.json 2> $O/${TAG}_bench_${q}_n1.err || echo "bench $q failed"
  tail -c 200 $O/${TAG}_bench_${q}_n1.json; echo
done
