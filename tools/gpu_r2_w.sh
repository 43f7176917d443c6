#!/bin/bash
mkdir -p gpurun_out
EGR_TC_GMAX_HALO=4 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2w_ops.log 2>&1; rc=$?; echo "ops exit $rc"; tail -n 3 gpurun_out/r2w_ops.log
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/op_times.py 1 > gpurun_out/r2w_ops_default.tsv 2>/dev/null
EGR_TC_GMAX_HALO=2 timeout 300 python tools/op_times.py 1 > gpurun_out/r2w_ops_h2.tsv 2>/dev/null
EGR_TC_GMAX_HALO=4 timeout 300 python tools/op_times.py 1 > gpurun_out/r2w_ops_h4.tsv 2>/dev/null
python tools/op_diff.py -v gpurun_out/r2w_ops_default.tsv gpurun_out/r2w_ops_h2.tsv gpurun_out/r2w_ops_h4.tsv
