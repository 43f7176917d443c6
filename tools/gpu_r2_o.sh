#!/bin/bash
# which operand stream limits the big non-halo GEMM?  (EGR_TC_DBG_SKIP: bit0 = no A loads, bit1 = no B loads)
mkdir -p gpurun_out
out=gpurun_out/r2o_skip.txt; : > $out
for mode in "EGR_TC_NO_PAIR=1" "EGR_TC_NO_PAIR=1 EGR_TC_NO_MT2=1" "EGR_TC_FORCE_PAIR=1"; do
  for skip in 0 1 2 3; do
    for sh in "conv2d 1024->1024 k3" "conv2d 512->512 k3" "conv2d 1024->1024 k1"; do
      echo "== $mode skip=$skip" >> $out
      env $mode EGR_TC_DBG_SKIP=$skip timeout 120 python tools/gemm_probe.py "$sh" 8 2>&1 | grep -v Warning | tail -n 1 >> $out
    done
  done
done
cat $out
