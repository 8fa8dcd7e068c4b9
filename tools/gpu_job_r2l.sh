#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu > $O/r2l_bench_$name.json 2> $O/r2l_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2l_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d["config"].get("eager",{})
    print("$name", "value %.1f img/s %.2f ms | eager %.2f graph %s | conv %s frac %.3f | fwd %.2f" % (d["value"], d["ms_per_step"], e.get("ms_per_step",0), e.get("graph_img_s"), r["ms_per_step"], r["frac"], d["config"]["fwd_loss_ms_per_step"]))
except Exception as ex:
    print("$name", "FAILED", ex)
PY
}
run side1 RYOLO_WGRAD_SIDE=1
run side0 RYOLO_WGRAD_SIDE=0
run side1b RYOLO_WGRAD_SIDE=1
run side0b RYOLO_WGRAD_SIDE=0
