#!/bin/bash
# closing evidence after the resident-weights change: tests, smoke, bench both arms, launch list with DRAM bytes
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl $O/model_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > $O/r4z2_pytest.log 2>&1; echo "rc=$?" >> $O/r4z2_pytest.log
tail -4 $O/r4z2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r4z2_smoke.log 2>&1; tail -1 $O/r4z2_smoke.log
timeout 900 python bench.py > $O/r4z2_bench.json 2> $O/r4z2_bench.err; tail -c 200 $O/r4z2_bench.json; tail -2 $O/r4z2_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r4z2_bench_ref.json 2> $O/r4z2_bench_ref.err; tail -c 200 $O/r4z2_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r4z2_launches_train.csv python tools/train_layers.py 32 > $O/r4z2_ncu_list.log 2>&1
ls -la $O | grep r4z2; date
