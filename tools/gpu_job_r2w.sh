#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward_ops.py tests/test_gpu_backward_model.py tests/test_config1.py -m gpu -q -p no:cacheprovider -x > $O/r2w_pytest.log 2>&1; echo "rc=$?" >> $O/r2w_pytest.log
tail -4 $O/r2w_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu > $O/r2w_bench_$i.json 2> $O/r2w_bench_$i.err
python - <<PY
import json
d=json.loads(open("$O/r2w_bench_$i.json").read().strip().splitlines()[-1])
print("run $i: %.1f img/s %.2f ms" % (d["value"], d["ms_per_step"]), d["config"]["eager"], d["roofline"]["serialized"]["ms_per_step"])
PY
done
