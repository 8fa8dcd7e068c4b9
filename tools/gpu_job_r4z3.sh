#!/bin/bash
# closing validation of the shipped defaults (kgrp 1, pair 256, wres 96, wg_x32 1): tests, smoke, bench both arms
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/config1_parity.jsonl $O/model_backward_parity.jsonl $O/model_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > $O/r4z3_pytest.log 2>&1; echo "rc=$?" >> $O/r4z3_pytest.log
tail -4 $O/r4z3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r4z3_smoke.log 2>&1; tail -1 $O/r4z3_smoke.log
timeout 900 python bench.py > $O/r4z3_bench.json 2> $O/r4z3_bench.err; tail -c 200 $O/r4z3_bench.json; tail -2 $O/r4z3_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r4z3_bench_ref.json 2> $O/r4z3_bench_ref.err; tail -c 200 $O/r4z3_bench_ref.json
date
