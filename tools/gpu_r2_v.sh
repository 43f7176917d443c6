#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2v_ops.log 2>&1; rc=$?; echo "ops+mega exit $rc"; tail -n 3 gpurun_out/r2v_ops.log
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/op_times.py 1 > gpurun_out/r2v_ops_default.tsv 2>/dev/null
EGR_TC_GMAX_HALO=4 timeout 300 python tools/op_times.py 1 > gpurun_out/r2v_ops_h4.tsv 2>/dev/null
python tools/op_diff.py -v gpurun_out/r2s_ops_old.tsv gpurun_out/r2v_ops_default.tsv gpurun_out/r2v_ops_h4.tsv
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -6
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -6
EGR_TC_GMAX=1 timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -6
