#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward_ops.py tests/test_gpu_backward_model.py -m gpu -q -p no:cacheprovider -x > $O/r2s_pytest.log 2>&1; echo "rc=$?" >> $O/r2s_pytest.log
tail -4 $O/r2s_pytest.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-aux --no-cpu > $O/r2s_bench_$name.json 2> $O/r2s_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2s_bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d["config"].get("eager",{})
    print("$name", "value %.1f img/s %.2f ms | eager %.2f graph %s | %s" % (d["value"], d["ms_per_step"], e.get("ms_per_step",0), e.get("graph_img_s"), r["serialized"]["ms_per_step"]))
except Exception as ex:
    print("$name", "FAILED", ex)
PY
  tail -2 $O/r2s_bench_$name.err
}
run fuse1 RYOLO_BN_FUSE=1
run fuse0 RYOLO_BN_FUSE=0
run fuse1b RYOLO_BN_FUSE=1
run fuse0b RYOLO_BN_FUSE=0
