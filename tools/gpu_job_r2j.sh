#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
RYOLO_BWD_PRIO=1 RYOLO_WG_BOXES=13 RYOLO_EW_REGS=2 timeout 300 python tools/timeline.py 32 0 400 > $O/r2j_timeline_full.txt 2>&1
RYOLO_BWD_PRIO=0 RYOLO_WG_BOXES=14 RYOLO_EW_REGS=0 timeout 300 python tools/timeline.py 32 0 400 > $O/r2j_timeline_base.txt 2>&1
head -3 $O/r2j_timeline_full.txt; head -3 $O/r2j_timeline_base.txt
sed -n 120,160p $O/r2j_timeline_full.txt
