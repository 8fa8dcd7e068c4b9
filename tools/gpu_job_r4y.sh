#!/bin/bash
# last check of the round: the whole GPU suite + smoke on the closing build
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r4y_pytest.log 2>&1; echo "rc=$?" >> $O/r4y_pytest.log
tail -3 $O/r4y_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r4y_smoke.log 2>&1; tail -1 $O/r4y_smoke.log
