#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl
date
timeout 1500 python -m pytest tests/test_config1.py tests/test_gpu_postprocess.py tests/test_gpu_backward_model.py -m gpu -q -p no:cacheprovider --durations=8 > $O/r2c_pytest.log 2>&1; echo "rc=$?" >> $O/r2c_pytest.log
tail -25 $O/r2c_pytest.log
date
timeout 300 python tools/debug_nms_adv.py > $O/r2c_nms_debug.log 2>&1; grep -c "equal" $O/r2c_nms_debug.log; grep "differ\|mismatching" $O/r2c_nms_debug.log | head
timeout 300 python tools/nms_bench.py 5 > $O/r2c_nms_bench.log 2>&1; tail -30 $O/r2c_nms_bench.log
date
