#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_postprocess.py tests/test_metrics.py tests/test_gpu_decode_loss.py -m gpu -q -p no:cacheprovider > $O/r2p_pytest.log 2>&1; echo "rc=$?" >> $O/r2p_pytest.log
tail -8 $O/r2p_pytest.log
for b in 1 0; do
  echo "== nms bench RYOLO_NMS_BAND=$b"
  RYOLO_NMS_BAND=$b timeout 600 python tools/nms_bench.py 5 > $O/r2p_nms_bench_band$b.log 2>&1; tail -4 $O/r2p_nms_bench_band$b.log | cut -c1-700
done
