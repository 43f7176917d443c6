#!/bin/bash
mkdir -p gpurun_out
for sk in 0 2 3; do
  EGR_TC_DBG_SKIP=$sk timeout 300 python tools/op_times.py 8 > gpurun_out/r2z_ops_b8_skip$sk.tsv 2>/dev/null
done
python tools/op_diff.py gpurun_out/r2z_ops_b8_skip*.tsv | head -24
