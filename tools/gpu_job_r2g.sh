#!/bin/bash
# elementwise rewrite: correctness of the new kernels, A/B micro-benchmark, bench
set +e
O=gpurun_out; mkdir -p $O
echo "== pytest (elementwise + model backward + graph)"; date
timeout 900 python -m pytest tests/test_gpu_backward_ops.py tests/test_gpu_conv.py tests/test_gpu_backward_model.py -m gpu -q -p no:cacheprovider -x > $O/r2g_pytest.log 2>&1; echo "rc=$?" >> $O/r2g_pytest.log
tail -8 $O/r2g_pytest.log
echo "== elementwise A/B"; date
timeout 600 python tools/elementwise_bench.py 32 > $O/r2g_elementwise.txt 2>&1; tail -22 $O/r2g_elementwise.txt
echo "== bench (no aux)"; date
timeout 600 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu > $O/r2g_bench.json 2> $O/r2g_bench.err; tail -c 600 $O/r2g_bench.json; tail -3 $O/r2g_bench.err
date
