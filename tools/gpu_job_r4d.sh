#!/bin/bash
# knock-out table of the conv pipeline on the warp-uniform build (which stage bounds the N <= 128 layers now?)
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python tools/diag_knobs.py 32 base,nomma,noaload,nobload,noload,noaload_nomma,nostore_nostats,halo > $O/r4d_diag.log 2>&1; tail -9 $O/r4d_diag.log
cp $O/diag_knobs_bs32.txt $O/r4d_diag_knobs_bs32.txt
