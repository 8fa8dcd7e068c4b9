#!/bin/bash
# Round-end evidence: tests, bench (both arms), v7 bench, launch list of the bench command, ncu full of the dominant kernels.
set +e
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; date
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
echo "== smoke"; date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench"; date
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json; tail -3 $O/bench.err
echo "== bench reference arm"; date
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 400 $O/bench_ref.json
echo "== bench v7"; date
timeout 600 python bench.py --workload train_v7 --no-cpu > $O/bench_v7.json 2> $O/bench_v7.err; tail -c 300 $O/bench_v7.json
echo "== nms / kfloss bench"; date
timeout 300 python tools/nms_bench.py 5 > $O/nms_bench.log 2>&1; tail -12 $O/nms_bench.log
echo "== ncu launch list of the bench command"; date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 760 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_list.log 2>&1
tail -2 $O/ncu_list.log | cut -c1-300
echo "== ncu dram bytes of every conv_fwd_kernel launch of one step"; date
timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_fwd_kernel -s 480 -c 240 --csv --log-file $O/conv_dram.csv \
   python tools/train_layers.py 32 > $O/ncu_dram.log 2>&1
tail -1 $O/ncu_dram.log | cut -c1-200
echo "== ncu full: conv fwd layers 17-24 (bs=32)"; date
timeout 420 ncu --set full --clock-control none -k regex:conv_fwd_kernel -s 237 -c 8 -o $O/conv_final python tools/conv_layers.py 32 yolov4 1 > $O/ncu_conv.log 2>&1
echo "== ncu full: wgrad (bs=32) mid + last"; date
timeout 420 ncu --set full --clock-control none -k regex:conv_wgrad_kernel -s 300 -c 6 -o $O/wgrad_final python tools/train_layers.py 32 > $O/ncu_wgrad.log 2>&1
for r in conv_final wgrad_final; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
  fi
done
ls -la $O | head -60; du -sh $O
date
