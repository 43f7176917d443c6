#!/bin/bash
for v in "EGR_TC_NO_PAIR=1" "EGR_TC_FORCE_PAIR=1" "EGR_TC_FORCE_PAIR=1 EGR_TC_PAIR_BN=128" "EGR_TC_FORCE_PAIR=1 EGR_TC_PAIR_BN=64" "EGR_TC_FORCE_PAIR=1 EGR_TC_STAGES=3" "EGR_TC_FORCE_PAIR=1 EGR_TC_STAGES=2"; do
  echo "=== $v"
  env $v timeout 200 python tools/gemm_probe.py "conv2d 1024->1024 k3" 8 2>&1 | grep -v Warning | tail -n 1
  env $v timeout 200 python tools/gemm_probe.py "conv2d 256->256 k3 d1 (256" 8 2>&1 | grep -v Warning | tail -n 1
done
