#!/bin/bash
set +e
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2 3; do
  timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/pytest_rep$i.log 2>&1; echo "rep $i rc=$?"; tail -2 $O/pytest_rep$i.log | head -1
done
echo "== memcheck (conv + wgrad + bn small tests)"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_conv.py tests/test_gpu_backward_ops.py -m gpu -q -x -k "shape0 or shape5 or shape8 or fused_bn or stem or bn_act" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "ERROR SUMMARY" $O/memcheck.log; tail -4 $O/memcheck.log
echo "== racecheck (conv epilogue staging)"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "test_conv_raw and shape1 and tma" > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/racecheck.log
