#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
rm -f $O/model_backward_parity.jsonl
date
timeout 1200 python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_decode_loss.py tests/test_gpu_model.py "tests/test_gpu_backward_model.py::test_whole_model_backward_vs_oracle_autograd" -m gpu -q -p no:cacheprovider > $O/r2d_pytest.log 2>&1; echo "rc=$?" >> $O/r2d_pytest.log
tail -12 $O/r2d_pytest.log
date
timeout 300 python tools/nms_bench.py 5 > $O/r2d_nms_bench.log 2>&1; tail -32 $O/r2d_nms_bench.log
date
echo "== ncu launch list of the non-conv kernels"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2d_nonconv_launches.csv python tools/nonconv_profile.py 2 > $O/r2d_ncu_list.log 2>&1
tail -2 $O/r2d_ncu_list.log | cut -c1-200
wc -l $O/r2d_nonconv_launches.csv
date
echo "== ncu full of the hot non-conv kernels (second repetition only)"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:kfloss_pairs|csl_pos|kf_pos|obj_dense|pp_score|pp_select|nms_mask|nms_scan|decode_|pairwise_iou' -o $O/r2d_nonconv python tools/nonconv_profile.py 1 > $O/r2d_ncu_full.log 2>&1
tail -3 $O/r2d_ncu_full.log | cut -c1-200
if [ -f $O/r2d_nonconv.ncu-rep ]; then ncu -i $O/r2d_nonconv.ncu-rep --page raw --csv > $O/r2d_nonconv_raw.csv 2>/dev/null; ls -la $O/r2d_nonconv*; fi
date
