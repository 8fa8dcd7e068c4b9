#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python tools/diag_knobs.py 32 base,noaload,nobload,noload > $O/r3g_diag.log 2>&1; tail -5 $O/r3g_diag.log
cp $O/diag_knobs_bs32.txt $O/r3g_diag_knobs_bs32.txt
