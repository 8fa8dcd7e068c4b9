#!/bin/bash
# round 2, call A: full GPU test suite (new parity tests) + baseline bench line
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl $O/model_parity.jsonl
nproc; free -g | head -2
date
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > $O/r2a_pytest.log 2>&1; echo "rc=$?" >> $O/r2a_pytest.log
tail -40 $O/r2a_pytest.log
date
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2a_bench.json 2> $O/r2a_bench.err; tail -c 1500 $O/r2a_bench.json; tail -3 $O/r2a_bench.err
date
