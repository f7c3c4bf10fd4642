#!/bin/bash
# 2-GPU refresh of the sharded bench lines (gpurun --gpus 2)
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --query q1 --e2e-steps 2 > gpurun_out/r01_bench_q1_n2.json 2> gpurun_out/r01_bench_q1_n2.err; tail -c 300 gpurun_out/r01_bench_q1_n2.json; tail -2 gpurun_out/r01_bench_q1_n2.err
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --query q3 --no-cpu --no-e2e > gpurun_out/r01_bench_q3_n2.json 2> gpurun_out/r01_bench_q3_n2.err; tail -c 300 gpurun_out/r01_bench_q3_n2.json; tail -2 gpurun_out/r01_bench_q3_n2.err
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --query groupby --no-e2e --no-cpu > gpurun_out/r01_bench_groupby_n2.json 2> gpurun_out/r01_bench_groupby_n2.err; tail -c 300 gpurun_out/r01_bench_groupby_n2.json; tail -2 gpurun_out/r01_bench_groupby_n2.err
