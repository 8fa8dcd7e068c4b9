#!/bin/bash
# kgrp (K blocks per ring stage, compile-time instantiations) A/B: conv parity tests, then per-layer timing tables
set +e
O=gpurun_out; mkdir -p $O
for g in 0 2 3; do
  RYOLO_KGRP=$g timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py tests/test_gpu_backward_ops.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
done
timeout 600 python tools/diag_knobs.py 32 base,kgrp1,kgrp2,kgrp3,nomma > $O/r4b_diag.log 2>&1; tail -6 $O/r4b_diag.log
cp $O/diag_knobs_bs32.txt $O/r4b_diag_knobs_bs32.txt
