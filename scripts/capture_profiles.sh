#!/bin/bash
# One-GPU capture of everything profiles/ cites: plain bench lines first (never under a profiler), then the ncu launch
# lists (--metrics gpu__time_duration.sum) and one --set full capture of the dominant kernels.  Run under gpurun.
TAG=${1:-r01}
O=gpurun_out
for q in q1 q6 q3 groupby; do
  extra=""; [ $q == groupby ] && extra="--no-e2e --no-cpu --steps 5 --warmup 4"; [ $q == q3 ] && extra="--no-cpu --steps 20"; [ $q == q6 ] && extra="--steps 20"
  timeout 600 python bench.py --query $q $extra > $O/${TAG}_bench_${q}_n1.json 2> $O/${TAG}_bench_${q}_n1.err || echo "bench $q failed"
  tail -c 300 $O/${TAG}_bench_${q}_n1.json; echo
done
NCU="ncu --clock-control none"
for q in q1 q3 groupby; do
  args="--query $q --steps 2 --warmup 1 --no-e2e --no-cpu"; [ $q == groupby ] && args="$args --rows 200000000 --groups 20000000"
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${TAG}_launches_${q}.csv python bench.py $args > $O/ncu_l_$q.log 2>&1
done
timeout 600 $NCU --set full --import-source on -k regex:k_fused_scan_agg_spec -c 1 -o $O/${TAG}_q1_fused_spec python bench.py --query q1 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f_q1.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_fused_scan_agg -c 2 -o $O/${TAG}_q3_probe_emit python bench.py --query q3 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f_q3.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_radix -c 8 -o $O/${TAG}_groupby_radix python bench.py --query groupby --rows 200000000 --groups 20000000 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f_gb.log 2>&1
tail -2 $O/ncu_f_gb.log
ls -la $O | grep "${TAG}_" | head -30
