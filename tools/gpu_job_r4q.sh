#!/bin/bash
# wgrad: 32-channel X boxes for Cin == 32 layers (knob wg_x32): parity tests with it on, then per-layer A/B
set +e
O=gpurun_out; mkdir -p $O
RYOLO_WG_X32=1 timeout 500 python -m pytest tests/test_gpu_backward_ops.py tests/test_gpu_bench_shapes.py tests/test_gpu_backward_model.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
timeout 400 python tools/diag_knobs.py 32 base,wg_x32,wg_x64 > $O/r4q_diag.log 2>&1; tail -4 $O/r4q_diag.log
cp $O/diag_knobs_bs32.txt $O/r4q_diag_knobs_bs32.txt
grep "^wgrad" $O/r4q_diag_knobs_bs32.txt | awk '$3<=64 && $2>=1280000'
