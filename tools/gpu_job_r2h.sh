#!/bin/bash
# overlap experiment: wgrad under the BatchNorm backward (stream priorities, 13-box wgrad, 96-register elementwise blocks)
set +e
O=gpurun_out; mkdir -p $O
echo "== pytest subset"; date
timeout 900 python -m pytest tests/test_gpu_backward_ops.py tests/test_gpu_backward_model.py -m gpu -q -p no:cacheprovider -x > $O/r2h_pytest.log 2>&1; echo "rc=$?" >> $O/r2h_pytest.log
tail -5 $O/r2h_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu > $O/r2h_bench_$name.json 2> $O/r2h_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2h_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d["config"].get("eager",{})
    print("$name", "value %.1f img/s %.2f ms | eager %.2f graph %s | conv %s frac %.3f" % (d["value"], d["ms_per_step"], e.get("ms_per_step",0), e.get("graph_img_s"), r["ms_per_step"], r["frac"]))
except Exception as ex:
    print("$name", "FAILED", ex)
PY
}
echo "== bench matrix"; date
run base RYOLO_BWD_PRIO=0 RYOLO_WG_BOXES=14 RYOLO_EW_REGS=0
run prio RYOLO_BWD_PRIO=1 RYOLO_WG_BOXES=14 RYOLO_EW_REGS=0
run b13 RYOLO_BWD_PRIO=0 RYOLO_WG_BOXES=13 RYOLO_EW_REGS=1
run full RYOLO_BWD_PRIO=1 RYOLO_WG_BOXES=13 RYOLO_EW_REGS=1
run full12 RYOLO_BWD_PRIO=1 RYOLO_WG_BOXES=12 RYOLO_EW_REGS=1
run full_nopdl RYOLO_BWD_PRIO=1 RYOLO_WG_BOXES=13 RYOLO_EW_REGS=1 RYOLO_PDL=0
date
