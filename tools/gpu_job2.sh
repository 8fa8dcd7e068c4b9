#!/bin/bash
# GPU call 2: validate the new defaults, bench, knob A/B, launch list, ncu full of the changed kernels.
set +e
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; date
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
echo "== bench"; date
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 1800 $O/bench.json; tail -3 $O/bench.err
echo "== diag knobs"; date
timeout 420 python tools/diag_knobs.py 32 > $O/diag_knobs.log 2>&1; tail -20 $O/diag_knobs.log
echo "== bench v7"; date
timeout 600 python bench.py --workload train_v7 --no-cpu > $O/bench_v7.json 2> $O/bench_v7.err; tail -c 900 $O/bench_v7.json
echo "== ncu launch list (train step)"; date
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 1480 -c 760 --csv --log-file $O/launches_train.csv \
   python tools/train_layers.py 32 > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log
echo "== ncu full: conv fwd layers 7-12 + 20,38 (bs=8)"; date
timeout 420 ncu --set full --clock-control none -k regex:conv_fwd_kernel -s 227 -c 6 -o $O/conv_tma python tools/conv_layers.py 8 yolov4 1 > $O/ncu_conv.log 2>&1
echo "== ncu full: wgrad last 8 (bs=8)"; date
timeout 420 ncu --set full --clock-control none -k regex:conv_wgrad_kernel -s 322 -c 8 -o $O/wgrad_hires python tools/train_layers.py 8 > $O/ncu_wgrad.log 2>&1
echo "== ncu full: bn bwd last 6 (bs=8)"; date
timeout 300 ncu --set full --clock-control none -k regex:bn_act_bwd -s 630 -c 6 -o $O/bn_hires python tools/train_layers.py 8 > $O/ncu_bn.log 2>&1
for r in conv_tma wgrad_hires bn_hires; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
    ncu -i $O/$r.ncu-rep --page source --csv > $O/${r}_source.csv 2>/dev/null
    gzip -f $O/${r}_source.csv
  fi
done
ls -la $O | head -50; du -sh $O
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/conv_tma.ncu-rep; fi
date
