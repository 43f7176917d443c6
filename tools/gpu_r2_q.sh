#!/bin/bash
# k-step period in CYCLES (trace build) with / without operand loads, pair and single-CTA tiles, batch 8
mkdir -p gpurun_out
export EGREGORA_B200_LIB=$PWD/comfyui-egregora-audio-super-resolution_b200/libegregora_b200_trace.so
out=gpurun_out/r2q_trace.txt; : > $out
for mode in "EGR_TC_NO_PAIR=1" "EGR_TC_FORCE_PAIR=1"; do
  for skip in 0 3 1 2; do
    echo "== $mode skip=$skip" >> $out
    env $mode EGR_TC_DBG_SKIP=$skip timeout 120 python tools/gemm_trace.py "conv2d 1024->1024 k3" 8 brief 2>&1 | grep -v Warning >> $out
  done
done
echo "== full trace pair" >> $out
EGR_TC_FORCE_PAIR=1 timeout 120 python tools/gemm_trace.py "conv2d 1024->1024 k3" 8 2>&1 | grep -v Warning >> $out
echo "== full trace pair skip 3" >> $out
EGR_TC_FORCE_PAIR=1 EGR_TC_DBG_SKIP=3 timeout 120 python tools/gemm_trace.py "conv2d 1024->1024 k3" 8 2>&1 | grep -v Warning >> $out
cat $out | cut -c 1-1500
