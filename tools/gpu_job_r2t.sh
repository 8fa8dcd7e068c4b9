#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python tools/diag_knobs.py 32 base,nostore,nostats,nostore_nostats,nomma,noaload,noaload_nostats,noaload_nomma,halo,acc2,nopdl > $O/r2t_diag.log 2>&1
tail -14 $O/r2t_diag.log
cp $O/diag_knobs_bs32.txt $O/r2t_diag_knobs_bs32.txt
