#!/bin/bash
mkdir -p gpurun_out
EGR_TC_HGROUP=2 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2y_ops.log 2>&1; rc=$?; echo "ops (hgroup 2) exit $rc"; tail -n 3 gpurun_out/r2y_ops.log
[ $rc -ne 0 ] && exit 1
EGR_TC_HGROUP=3 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2y_ops3.log 2>&1; rc=$?; echo "ops (hgroup 3) exit $rc"; tail -n 3 gpurun_out/r2y_ops3.log
[ $rc -ne 0 ] && exit 1
for g in 1 2 3 4; do
  EGR_TC_HGROUP=$g timeout 300 python tools/op_times.py 1 > gpurun_out/r2y_ops_hg$g.tsv 2>/dev/null
done
python tools/op_diff.py -v gpurun_out/r2y_ops_hg*.tsv
for g in 1 2 4; do
  EGR_TC_HGROUP=$g timeout 300 python tools/op_times.py 8 > gpurun_out/r2y_ops_b8_hg$g.tsv 2>/dev/null
done
python tools/op_diff.py -v gpurun_out/r2y_ops_b8_hg*.tsv
