#!/bin/bash
# Round-2 ncu evidence (one GPU, under gpurun): launch lists of the per-step kernels and --set full captures of the dominant
# kernels -- each only after the same command has run without a profiler.  Summaries: python scripts/summarize_ncu.py r02
TAG=r02
O=gpurun_out
NCU="ncu --clock-control none"
for q in q1 q3; do
  args="--query $q --steps 2 --warmup 1 --no-e2e --no-cpu --extra-queries="
  QGPU_BENCH_SKIP_GENERIC=1 timeout 300 python bench.py $args > $O/${TAG}_plain_$q.json 2> $O/${TAG}_plain_$q.err || echo "plain $q failed"
  QGPU_BENCH_SKIP_GENERIC=1 timeout 600 $NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file $O/${TAG}_launches_${q}.csv python bench.py $args > $O/ncu_l_$q.log 2>&1
done
A="--steps 1 --warmup 0 --no-e2e --no-cpu --extra-queries="
QGPU_BENCH_SKIP_GENERIC=1 timeout 600 $NCU --set full --import-source on -k regex:k_fused_scan_agg_spec -c 1 -o $O/${TAG}_q1_fused_spec python bench.py --query q1 $A > $O/ncu_f_q1.log 2>&1
QGPU_BENCH_SKIP_GENERIC=1 QGPU_FUSED_NOAOT=1 timeout 600 $NCU --set full --import-source on -k regex:k_fused_scan_agg_jit -c 1 -o $O/${TAG}_q1_fused_jit python bench.py --query q1 $A > $O/ncu_f_q1j.log 2>&1
QGPU_BENCH_SKIP_GENERIC=1 timeout 600 $NCU --set full --import-source on -k regex:k_fused_scan_agg -c 3 -o $O/${TAG}_q3_build_emit_probe python bench.py --query q3 $A > $O/ncu_f_q3.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_csv -c 4 -o $O/${TAG}_csv_parse python scripts/csv_probe.py 1 > $O/ncu_f_csv.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_dict_insert -c 1 -o $O/${TAG}_dict_insert python bench.py --query q1 $A > $O/ncu_f_dict.log 2>&1
ls -la $O | grep "${TAG}_" | tail -20
