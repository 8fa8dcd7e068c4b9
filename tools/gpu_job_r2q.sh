#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_postprocess.py -m gpu -q -p no:cacheprovider > $O/r2q_pytest.log 2>&1; echo "rc=$?" >> $O/r2q_pytest.log
tail -3 $O/r2q_pytest.log
for b in 4 8 16; do
  RYOLO_NMS_BAND=$b timeout 600 python tools/nms_bench.py 5 > $O/r2q_nms_bench_band$b.log 2>&1
  echo "band=$b $(grep -A1 '"nc' $O/r2q_nms_bench_band$b.log | grep ms | tr -d ' \n')"
done
