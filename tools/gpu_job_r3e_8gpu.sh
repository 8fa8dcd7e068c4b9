#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
date; nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 --no-aux > $O/r3e_bench_8gpu.json 2> $O/r3e_bench_8gpu.err; tail -c 600 $O/r3e_bench_8gpu.json; tail -3 $O/r3e_bench_8gpu.err
date
date
