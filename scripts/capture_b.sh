#!/bin/bash
# partial refresh (one GPU): group-by bench line at full size, ncu of the radix kernels and of the Q6 scan kernel
TAG=${1:-r01}
O=gpurun_out
timeout 600 python bench.py --query groupby --no-e2e --no-cpu --steps 5 --warmup 4 > $O/${TAG}_bench_groupby_n1.json 2> $O/${TAG}_bench_groupby_n1.err || echo "bench groupby failed"
tail -c 600 $O/${TAG}_bench_groupby_n1.json; echo
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${TAG}_launches_groupby.csv python bench.py --query groupby --steps 2 --warmup 1 --no-e2e --no-cpu --rows 200000000 --groups 20000000 > $O/ncu_l_groupby.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_radix -c 8 -o $O/${TAG}_groupby_radix -f python bench.py --query groupby --rows 200000000 --groups 20000000 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f_gb.log 2>&1
tail -2 $O/ncu_f_gb.log
timeout 600 $NCU --set full --import-source on -k regex:k_fused_scan_agg_spec -c 1 -o $O/${TAG}_q6_fused_spec -f python bench.py --query q6 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_f_q6.log 2>&1
tail -2 $O/ncu_f_q6.log
