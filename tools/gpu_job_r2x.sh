#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
for i in 1 2 3; do for f in 1 0; do
RYOLO_RES_FUSE=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu --no-graph > $O/r2x_bench.json 2> $O/r2x_bench.err
python - <<PY
import json
d=json.loads(open("$O/r2x_bench.json").read().strip().splitlines()[-1])
print("fuse=$f run $i: %.1f img/s %.2f ms" % (d["value"], d["ms_per_step"]), d["roofline"]["serialized"]["ms_per_step"])
PY
done; done
