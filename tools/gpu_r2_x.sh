#!/bin/bash
mkdir -p gpurun_out
for b in 2 4 8 16; do
  timeout 300 python tools/op_times.py $b > gpurun_out/r2x_ops_b${b}_h1.tsv 2>/dev/null
  EGR_TC_GMAX_HALO=4 timeout 300 python tools/op_times.py $b > gpurun_out/r2x_ops_b${b}_h4.tsv 2>/dev/null
  python tools/op_diff.py -v gpurun_out/r2x_ops_b${b}_h1.tsv gpurun_out/r2x_ops_b${b}_h4.tsv
done
