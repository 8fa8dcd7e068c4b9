#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
RYOLO_BN_DEFER=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r3b_launches_train.csv python tools/train_layers.py 32 > $O/r3b_ncu_list.log 2>&1
python tools/agg_launches.py $O/r3b_launches_train.csv --steps=3 | head -16
