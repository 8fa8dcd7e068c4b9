#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl
date
timeout 600 python tools/debug_nms_adv.py > $O/r2b_nms_debug.log 2>&1; tail -60 $O/r2b_nms_debug.log
date
timeout 1500 python -m pytest tests/test_config1.py tests/test_gpu_backward_model.py -m gpu -q -p no:cacheprovider --durations=8 > $O/r2b_pytest.log 2>&1; echo "rc=$?" >> $O/r2b_pytest.log
tail -25 $O/r2b_pytest.log
date
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err; tail -c 4000 $O/r2b_bench.json; tail -5 $O/r2b_bench.err
date
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_bench_ref.json 2> $O/r2b_bench_ref.err; tail -c 600 $O/r2b_bench_ref.json
date
