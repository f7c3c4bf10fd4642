#!/bin/bash
# 2-GPU confirmation run (gpurun --gpus 2): repartition tests, then the sharded benches
set -x
N=${1:-2}
python -m pytest tests/test_gpu_repartition.py -x -q > gpurun_out/r01_pytest_repart.log 2>&1; tail -3 gpurun_out/r01_pytest_repart.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --query q1 > gpurun_out/bench_q1_n$N.json 2> gpurun_out/bench_q1_n$N.err; tail -c 400 gpurun_out/bench_q1_n$N.json; tail -3 gpurun_out/bench_q1_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --query q3 --no-cpu > gpurun_out/bench_q3_n$N.json 2> gpurun_out/bench_q3_n$N.err; tail -c 400 gpurun_out/bench_q3_n$N.json; tail -3 gpurun_out/bench_q3_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --query groupby --no-e2e --no-cpu > gpurun_out/bench_groupby_n$N.json 2> gpurun_out/bench_groupby_n$N.err; tail -c 400 gpurun_out/bench_groupby_n$N.json; tail -3 gpurun_out/bench_groupby_n$N.err
