#!/bin/bash
set +e
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; date
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
echo "== bench"; date
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 1800 $O/bench.json; tail -3 $O/bench.err
echo "== diag knobs"; date
timeout 420 python tools/diag_knobs.py 32 base,wg_uniform,wg_tap1,wg_nomma,wg_noload,bn_old > $O/diag_knobs.log 2>&1; tail -8 $O/diag_knobs.log
echo "== nms / kfloss bench"; date
timeout 300 python tools/nms_bench.py 5 > $O/nms_bench.log 2>&1; tail -30 $O/nms_bench.log
echo "== ncu launch list (train step)"; date
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 1480 -c 760 --csv --log-file $O/launches_train.csv \
   python tools/train_layers.py 32 > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log
date
