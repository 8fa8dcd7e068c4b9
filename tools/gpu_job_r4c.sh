#!/bin/bash
# warp-uniform role dispatch (warp index through a shuffle) A/B: parity tests on the new build, then alternating
# per-layer timing runs of the two libraries (U = uniform, N = the build before the change) on one box
set +e
O=gpurun_out; mkdir -p $O
L=r-yolov4_b200
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py tests/test_gpu_backward_ops.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
cp $L/libryolo_b200.so /tmp/U.so; cp $L/libryolo_b200_nouni.so /tmp/N.so
for v in U N U N; do
  cp /tmp/$v.so $L/libryolo_b200.so
  timeout 300 python tools/diag_knobs.py 32 base > $O/r4c_diag_$v.log 2>&1; echo $v $(tail -1 $O/r4c_diag_$v.log)
  cp $O/diag_knobs_bs32.txt $O/r4c_diag_knobs_bs32_$v.txt
done
cp /tmp/U.so $L/libryolo_b200.so
