#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
for h in 1 2; do
  echo "== conv tests with RYOLO_HALO=$h"
  RYOLO_HALO=$h timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py tests/test_gpu_backward_ops.py -m gpu -q -p no:cacheprovider > $O/r2n_pytest_halo$h.log 2>&1
  tail -4 $O/r2n_pytest_halo$h.log
  grep -c "^FAILED" $O/r2n_pytest_halo$h.log
done
for h in 0 2; do
  echo "== conv layers RYOLO_HALO=$h"
  RYOLO_HALO=$h timeout 300 python tools/conv_layers.py 32 yolov4 3 2>&1 | grep -E "K= 1152|K=  576|K=  288" | head -12
done
