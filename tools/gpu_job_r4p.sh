#!/bin/bash
# resident weights (knob wres): parity tests with it on, then per-layer A/B
set +e
O=gpurun_out; mkdir -p $O
RYOLO_WRES=128 timeout 500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py tests/test_gpu_backward_ops.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
timeout 400 python tools/diag_knobs.py 32 base,wres48,wres96,wres128,wres0 > $O/r4p_diag.log 2>&1; tail -6 $O/r4p_diag.log
cp $O/diag_knobs_bs32.txt $O/r4p_diag_knobs_bs32.txt
