#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2m_pytest.log 2>&1; echo "rc=$?" >> $O/r2m_pytest.log
tail -6 $O/r2m_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2m_bench.json 2> $O/r2m_bench.err
python - <<PY
import json
d=json.loads(open("$O/r2m_bench.json").read().strip().splitlines()[-1])
print("v4 %.1f img/s %.2f ms" % (d["value"], d["ms_per_step"]), d["config"].get("eager"))
a=d["aux"]["train_v7"]; print("v7 %.1f img/s %.2f ms" % (a["value"], a["ms_per_step"]), a.get("eager"))
PY
tail -3 $O/r2m_bench.err
