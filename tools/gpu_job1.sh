#!/bin/bash
# GPU call 1: validate, bench, knob A/B, ncu launch list + full captures.  Every step has its own timeout.
set +e
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
echo "== pytest (baseline paths)"; date
timeout 900 python -m pytest tests -m gpu -q -k "not tma" -p no:cacheprovider > $O/pytest_base.log 2>&1; echo "rc=$?" >> $O/pytest_base.log
tail -3 $O/pytest_base.log
echo "== pytest (tma epilogue)"; date
timeout 300 python -m pytest tests -m gpu -q -k "tma" -p no:cacheprovider > $O/pytest_tma.log 2>&1; echo "rc=$?" >> $O/pytest_tma.log
tail -3 $O/pytest_tma.log
echo "== bench"; date
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json
echo "== diag knobs"; date
timeout 420 python tools/diag_knobs.py 32 > $O/diag_knobs.log 2>&1; cat $O/diag_knobs.log | tail -20
echo "== ncu launch list (train step)"; date
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 1480 -c 760 --csv --log-file $O/launches_train.csv \
   python tools/train_layers.py 32 > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log
echo "== ncu full: conv fwd layers 0-8 (bs=8)"; date
timeout 420 ncu --set full --clock-control none -k regex:conv_fwd_kernel -s 220 -c 9 -o $O/conv_hires python tools/conv_layers.py 8 yolov4 1 > $O/ncu_conv.log 2>&1
echo "== ncu full: wgrad last 8 (bs=8)"; date
timeout 420 ncu --set full --clock-control none -k regex:conv_wgrad_kernel -s 322 -c 8 -o $O/wgrad_hires python tools/train_layers.py 8 > $O/ncu_wgrad.log 2>&1
echo "== ncu full: bn bwd last 6 (bs=8)"; date
timeout 300 ncu --set full --clock-control none -k regex:bn_act_bwd -s 630 -c 6 -o $O/bn_hires python tools/train_layers.py 8 > $O/ncu_bn.log 2>&1
for r in conv_hires wgrad_hires bn_hires; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
    ncu -i $O/$r.ncu-rep --page source --csv > $O/${r}_source.csv 2>/dev/null
    gzip -f $O/${r}_source.csv
  fi
done
ls -la $O; du -sh $O
# keep the merge under the 64 MiB cap
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/conv_hires.ncu-rep; fi
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/wgrad_hires.ncu-rep; fi
date
