#!/bin/bash
# Round-2 refresh after the radix group-by rework (one GPU, under gpurun): the default bench line, the group-by line at full
# size, then -- each only after the same command ran without a profiler -- the launch list of a group-by step and --set full
# captures of the radix kernels at 200 M rows / 20 M groups (QGPU_RADIX_B2=9: the level-2 fan-out of the full-size run).
TAG=r02b
O=gpurun_out
NCU="ncu --clock-control none"
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err || echo "default bench failed"
tail -c 300 $O/${TAG}_bench_n1.json; echo
timeout 600 python bench.py --query groupby --no-e2e --steps 5 --warmup 3 --extra-queries "" > $O/${TAG}_bench_groupby_n1.json 2> $O/${TAG}_bench_groupby_n1.err || echo "bench groupby failed"
tail -c 300 $O/${TAG}_bench_groupby_n1.json; echo
G="--query groupby --rows 200000000 --groups 20000000 --no-e2e --no-cpu --extra-queries="
QGPU_RADIX_B2=9 timeout 300 python bench.py $G --steps 2 --warmup 1 > $O/${TAG}_plain_groupby_200m.json 2> $O/${TAG}_plain_groupby_200m.err || echo "plain failed"
QGPU_RADIX_B2=9 timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${TAG}_launches_groupby.csv python bench.py $G --steps 2 --warmup 1 > $O/ncu_l_groupby.log 2>&1
QGPU_RADIX_B2=9 timeout 600 $NCU --set full --import-source on -k regex:k_radix -c 7 -o $O/${TAG}_groupby_radix -f python bench.py $G --steps 1 --warmup 0 > $O/ncu_f_gb.log 2>&1
tail -2 $O/ncu_f_gb.log | cut -c1-200
ls -la $O | grep "${TAG}_" | tail -20
