#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2u_ops.log 2>&1; rc=$?; echo "ops exit $rc"; tail -n 3 gpurun_out/r2u_ops.log
[ $rc -ne 0 ] && exit 1
for cfg in "1 1" "4 2" "4 4"; do
  set -- $cfg
  EGR_TC_GMAX=$1 EGR_TC_GMAX_HALO=$2 timeout 300 python tools/op_times.py 1 > gpurun_out/r2u_ops_g$1_h$2.tsv 2>/dev/null
done
python tools/op_diff.py -v gpurun_out/r2u_ops_g*.tsv
