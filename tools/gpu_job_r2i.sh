#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
for cfg in "RYOLO_WG_BOXES=14 RYOLO_EW_REGS=0" "RYOLO_WG_BOXES=13 RYOLO_EW_REGS=2" "RYOLO_WG_BOXES=13 RYOLO_EW_REGS=1" "RYOLO_WG_BOXES=12 RYOLO_EW_REGS=1"; do
  echo "== probe $cfg"
  env $cfg timeout 300 python tools/overlap_probe.py 32 2>&1 | tail -6
done | tee $O/r2i_overlap_probe.txt
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu > $O/r2i_bench_$name.json 2> $O/r2i_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2i_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d["config"].get("eager",{})
    print("$name", "value %.1f img/s %.2f ms | eager %.2f graph %s | conv %s frac %.3f" % (d["value"], d["ms_per_step"], e.get("ms_per_step",0), e.get("graph_img_s"), r["ms_per_step"], r["frac"]))
except Exception as ex:
    print("$name", "FAILED", ex)
PY
}
run base RYOLO_BWD_PRIO=0 RYOLO_WG_BOXES=14 RYOLO_EW_REGS=0
run full RYOLO_BWD_PRIO=1 RYOLO_WG_BOXES=13 RYOLO_EW_REGS=1
run full_noprio RYOLO_BWD_PRIO=0 RYOLO_WG_BOXES=13 RYOLO_EW_REGS=1
