#!/bin/bash
# round 2 evidence job: full GPU test suite, smoke, bench (both arms), launch list + DRAM bytes of one training step,
# ncu --set full of the conv / wgrad kernels
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl $O/model_parity.jsonl
echo "== pytest"; date
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > $O/r2u_pytest.log 2>&1; echo "rc=$?" >> $O/r2u_pytest.log
tail -22 $O/r2u_pytest.log
echo "== smoke"; date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2u_smoke.log 2>&1; tail -2 $O/r2u_smoke.log
echo "== bench"; date
timeout 900 python bench.py > $O/r2u_bench.json 2> $O/r2u_bench.err; tail -c 300 $O/r2u_bench.json; tail -3 $O/r2u_bench.err
echo "== bench reference arm"; date
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2u_bench_ref.json 2> $O/r2u_bench_ref.err; tail -c 300 $O/r2u_bench_ref.json
echo "== ncu launch list + DRAM bytes of one training step (third step of tools/train_layers.py)"; date
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2u_launches_train.csv python tools/train_layers.py 32 > $O/r2u_ncu_list.log 2>&1
tail -2 $O/r2u_ncu_list.log | cut -c1-200; wc -l $O/r2u_launches_train.csv
echo "== ncu full: conv fwd layers 17-24 (bs=32)"; date
timeout 420 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 237 -c 8 -o $O/r2u_conv python tools/conv_layers.py 32 yolov4 1 > $O/r2u_ncu_conv.log 2>&1
echo "== ncu full: wgrad (bs=32) mid + last"; date
timeout 420 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 300 -c 6 -o $O/r2u_wgrad python tools/train_layers.py 32 > $O/r2u_ncu_wgrad.log 2>&1
for r in r2u_conv r2u_wgrad; do
  if [ -f $O/$r.ncu-rep ]; then ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; rm -f $O/$r.ncu-rep; fi
done
ls -la $O | grep r2f; date
echo "== non-conv: nms bench, ncu launch list, ncu full"; date
timeout 300 python tools/nms_bench.py 5 > $O/r2u_nms_bench.log 2>&1; grep -A2 '"nc' $O/r2u_nms_bench.log | grep -E "nc|ms" | paste - - | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2u_nonconv_launches.csv python tools/nonconv_profile.py 2 > $O/r2u_ncu_list2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:kfloss_pairs|csl_pos|kfiou_pos|obj_dense|pp_score|pp_select|decode_|pairwise_iou' -o $O/r2u_nonconv python tools/nonconv_profile.py 1 > $O/r2u_ncu_full2.log 2>&1
if [ -f $O/r2u_nonconv.ncu-rep ]; then ncu -i $O/r2u_nonconv.ncu-rep --page raw --csv > $O/r2u_nonconv_raw.csv 2>/dev/null; rm -f $O/r2u_nonconv.ncu-rep; fi
ls -la $O | grep r2u; date
