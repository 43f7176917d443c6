#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2t_ops.log 2>&1; rc=$?; echo "ops exit $rc"; tail -n 3 gpurun_out/r2t_ops.log
[ $rc -ne 0 ] && exit 1
for cfg in "1 1" "4 1" "4 2" "4 4" "2 2" "3 3"; do
  set -- $cfg
  EGR_TC_GMAX=$1 EGR_TC_GMAX_HALO=$2 timeout 300 python tools/op_times.py 1 > gpurun_out/r2t_ops_g$1_h$2.tsv 2>/dev/null
done
EGR_TC_GMAX=4 EGR_TC_GMAX_HALO=4 timeout 300 python tools/op_times.py 8 > gpurun_out/r2t_ops_b8_g4_h4.tsv 2>/dev/null
EGR_TC_GMAX=4 EGR_TC_GMAX_HALO=1 timeout 300 python tools/op_times.py 8 > gpurun_out/r2t_ops_b8_g4_h1.tsv 2>/dev/null
EGR_TC_GMAX=1 EGR_TC_GMAX_HALO=1 timeout 300 python tools/op_times.py 8 > gpurun_out/r2t_ops_b8_g1_h1.tsv 2>/dev/null
wc -l gpurun_out/r2t_ops_*.tsv
