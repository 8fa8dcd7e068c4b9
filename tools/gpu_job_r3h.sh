#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
RYOLO_PAIR=1 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py tests/test_gpu_backward_ops.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/diag_knobs.py 32 base,pair > $O/r3h_diag.log 2>&1; tail -3 $O/r3h_diag.log
cp $O/diag_knobs_bs32.txt $O/r3h_diag_knobs_bs32.txt
