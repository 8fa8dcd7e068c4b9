#!/bin/bash
# round-2 closing evidence (after kgrp / warp-uniform dispatch / cluster pairs): tests, smoke, bench both arms,
# launch list with DRAM bytes, raw ncu pages of conv (incl. a paired deep layer) and wgrad (binary reports dropped after export)
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl $O/model_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > $O/r4z_pytest.log 2>&1; echo "rc=$?" >> $O/r4z_pytest.log
tail -4 $O/r4z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r4z_smoke.log 2>&1; tail -1 $O/r4z_smoke.log
timeout 900 python bench.py > $O/r4z_bench.json 2> $O/r4z_bench.err; tail -c 200 $O/r4z_bench.json; tail -2 $O/r4z_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r4z_bench_ref.json 2> $O/r4z_bench_ref.err; tail -c 200 $O/r4z_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r4z_launches_train.csv python tools/train_layers.py 32 > $O/r4z_ncu_list.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 237 -c 8 -o $O/r4z_conv python tools/conv_layers.py 32 yolov4 1 > $O/r4z_ncu_conv.log 2>&1
if [ -f $O/r4z_conv.ncu-rep ]; then ncu -i $O/r4z_conv.ncu-rep --page raw --csv > $O/r4z_conv_raw.csv 2>/dev/null; rm -f $O/r4z_conv.ncu-rep; fi
timeout 420 ncu --set full --clock-control none -k regex:conv_fwd_kernel -s 288 -c 8 -o $O/r4z_conv_deep python tools/conv_layers.py 32 yolov4 1 > $O/r4z_ncu_conv_deep.log 2>&1
if [ -f $O/r4z_conv_deep.ncu-rep ]; then ncu -i $O/r4z_conv_deep.ncu-rep --page raw --csv > $O/r4z_conv_deep_raw.csv 2>/dev/null; rm -f $O/r4z_conv_deep.ncu-rep; fi
timeout 420 ncu --set full --clock-control none -k regex:conv_wgrad_kernel -s 300 -c 6 -o $O/r4z_wgrad python tools/train_layers.py 32 > $O/r4z_ncu_wgrad.log 2>&1
if [ -f $O/r4z_wgrad.ncu-rep ]; then ncu -i $O/r4z_wgrad.ncu-rep --page raw --csv > $O/r4z_wgrad_raw.csv 2>/dev/null; rm -f $O/r4z_wgrad.ncu-rep; fi
ls -la $O | grep r4z; date
