#!/bin/bash
# 2-GPU check of the closing build: cluster-pair conv kernels next to NCCL's kernels, bucketed all-reduce
set +e
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_check.py > $O/r4n_ddp_check.log 2>&1; tail -4 $O/r4n_ddp_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/r4n_bench_2gpu.json 2> $O/r4n_bench_2gpu.err; tail -c 700 $O/r4n_bench_2gpu.json; tail -3 $O/r4n_bench_2gpu.err
