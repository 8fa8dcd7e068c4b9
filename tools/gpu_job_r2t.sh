#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python tools/diag_knobs.py 32 base,k32_n128,k32_all > $O/r2t_diag.log 2>&1
tail -14 $O/r2t_diag.log
cp $O/diag_knobs_bs32.txt $O/r2t_diag_knobs_bs32.txt
RYOLO_SW64=2 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward_ops.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
