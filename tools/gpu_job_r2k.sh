#!/bin/bash
# round 2 evidence job: full GPU test suite, smoke, bench (both arms), launch list + DRAM bytes of one training step,
# ncu --set full of the conv / wgrad kernels
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl $O/model_parity.jsonl
echo "== pytest"; date
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > $O/r2k_pytest.log 2>&1; echo "rc=$?" >> $O/r2k_pytest.log
tail -22 $O/r2k_pytest.log
echo "== smoke"; date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; tail -2 $O/r2k_smoke.log
echo "== bench"; date
timeout 900 python bench.py > $O/r2k_bench.json 2> $O/r2k_bench.err; tail -c 300 $O/r2k_bench.json; tail -3 $O/r2k_bench.err
echo "== bench reference arm"; date
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2k_bench_ref.json 2> $O/r2k_bench_ref.err; tail -c 300 $O/r2k_bench_ref.json
echo "== ncu launch list + DRAM bytes of one training step (third step of tools/train_layers.py)"; date
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2k_launches_train.csv python tools/train_layers.py 32 > $O/r2k_ncu_list.log 2>&1
tail -2 $O/r2k_ncu_list.log | cut -c1-200; wc -l $O/r2k_launches_train.csv
echo "== ncu full: conv fwd layers 17-24 (bs=32)"; date
timeout 420 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel -s 237 -c 8 -o $O/r2k_conv python tools/conv_layers.py 32 yolov4 1 > $O/r2k_ncu_conv.log 2>&1
echo "== ncu full: wgrad (bs=32) mid + last"; date
timeout 420 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 300 -c 6 -o $O/r2k_wgrad python tools/train_layers.py 32 > $O/r2k_ncu_wgrad.log 2>&1
for r in r2k_conv r2k_wgrad; do
  if [ -f $O/$r.ncu-rep ]; then ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; fi
done
ls -la $O | grep r2f; date
