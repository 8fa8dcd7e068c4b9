#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r3a_pytest.log 2>&1; echo "rc=$?" >> $O/r3a_pytest.log
tail -5 $O/r3a_pytest.log
for i in 1 2; do for f in 1 0; do
RYOLO_BN_DEFER=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu --no-graph > $O/r3a_bench.json 2> $O/r3a_bench.err
python - <<PY
import json
d=json.loads(open("$O/r3a_bench.json").read().strip().splitlines()[-1])
print("defer=$f run $i: %.1f img/s %.2f ms fwd+loss %.2f" % (d["value"], d["ms_per_step"], d["config"]["fwd_loss_ms_per_step"]), d["roofline"]["serialized"]["ms_per_step"])
PY
done; done
