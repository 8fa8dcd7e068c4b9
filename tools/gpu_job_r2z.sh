#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python tools/diag_knobs.py 32 base,nostats,nostatstail > $O/r2z_diag.log 2>&1; tail -4 $O/r2z_diag.log
cp $O/diag_knobs_bs32.txt $O/r2z_diag_knobs_bs32.txt
