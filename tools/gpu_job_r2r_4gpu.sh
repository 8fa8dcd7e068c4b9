#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
date; nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 10 --warmup 3 --no-aux > $O/r2r_bench_4gpu.json 2> $O/r2r_bench_4gpu.err; tail -c 600 $O/r2r_bench_4gpu.json; tail -3 $O/r2r_bench_4gpu.err
date
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-aux --no-cpu > $O/r2r_bench_1gpu.json 2> $O/r2r_bench_1gpu.err; tail -c 300 $O/r2r_bench_1gpu.json
date
