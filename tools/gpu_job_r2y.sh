#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
for i in 1 2; do for f in 1 0; do
RYOLO_WG_CLEAR=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu --no-graph > $O/r2y_bench.json 2> $O/r2y_bench.err
python - <<PY
import json
d=json.loads(open("$O/r2y_bench.json").read().strip().splitlines()[-1])
print("clear=$f run $i: %.1f img/s %.2f ms" % (d["value"], d["ms_per_step"]), d["roofline"]["serialized"]["ms_per_step"], d["clocks"]["sm_mhz"])
PY
done; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2y_launches.csv -k regex:unpack_wgrad python tools/train_layers.py 32 > /dev/null 2>&1
grep unpack $O/r2y_launches.csv | tail -3 | cut -c1-60,200-400
