#!/bin/bash
# cluster pairs with weight-tile multicast (knob pair): parity tests with pairing forced on every eligible layer, then A/B
set +e
O=gpurun_out; mkdir -p $O
RYOLO_PAIR=1056 timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py tests/test_gpu_backward_ops.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout 400 python tools/diag_knobs.py 32 base,pair256,pair128,pair64 > $O/r4e_diag.log 2>&1; tail -6 $O/r4e_diag.log
cp $O/diag_knobs_bs32.txt $O/r4e_diag_knobs_bs32.txt
