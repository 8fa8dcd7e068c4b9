#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
date; nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_check.py > $O/r2o_ddp_check.log 2>&1; tail -6 $O/r2o_ddp_check.log
date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2o_bench_2gpu.json 2> $O/r2o_bench_2gpu.err; tail -c 1500 $O/r2o_bench_2gpu.json; tail -3 $O/r2o_bench_2gpu.err
date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2o_bench_ref_2gpu.json 2> $O/r2o_bench_ref_2gpu.err; tail -c 400 $O/r2o_bench_ref_2gpu.json
date
