#!/bin/bash
set +e
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; date
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
echo "== diag knobs"; date
timeout 420 python tools/diag_knobs.py 32 base,nopdl,acc2,nostats,nostore > $O/diag_knobs.log 2>&1; tail -8 $O/diag_knobs.log
echo "== bench"; date
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 1800 $O/bench.json; tail -3 $O/bench.err
date
