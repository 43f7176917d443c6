#!/bin/bash
# last GPU seconds of the round: gn_fused with four loads in flight — op tests, timing of the small-map norms at batch 8, c2 parity
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "groupnorm" > gpurun_out/r2ss_ops.log 2>&1; rc=$?; echo "gn ops exit $rc"; tail -n 1 gpurun_out/r2ss_ops.log
[ $rc -ne 0 ] && exit 1
timeout 60 python tools/op_times.py 8 "mid.block_1.norm" 2>/dev/null | awk -F'\t' '{print $2, $8}'
timeout 60 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -p no:cacheprovider -x -k "c2_full_spec" > gpurun_out/r2ss_c2.log 2>&1; echo "c2 parity exit $?"; tail -n 1 gpurun_out/r2ss_c2.log
