#!/bin/bash
# memcheck of the operand-ring variants (cluster pairs incl. ghost tiles, grouped K blocks) on a few shapes
set +e
O=gpurun_out; mkdir -p $O
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "operand_ring and (shape0 or shape2 or shape6 or shape8 or shape11) and (pair_any or kgrp_wide)" > $O/r4m_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "ERROR SUMMARY: 0 errors" $O/r4m_memcheck.log; tail -4 $O/r4m_memcheck.log
