#!/bin/bash
mkdir -p gpurun_out
EGR_TC_FORCE_PAIR=1 timeout 400 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2k_ops_pair.log 2>&1; echo "ops (forced pair) exit $?"; tail -n 5 gpurun_out/r2k_ops_pair.log
for v in "EGR_TC_NO_PAIR=1" "EGR_TC_FORCE_PAIR=1" "EGR_TC_FORCE_PAIR=1 EGR_TC_ONE_ISSUER=1" "EGR_TC_FORCE_PAIR=1 EGR_TC_PAIR_BN=128"; do
  echo "=== $v"
  for sh in "conv2d 1024->1024 k3" "conv2d 256->256 k3 d1 (256" "conv2d 128->128 k3 d1 (512" "conv2d 512->512 k3"; do
    env $v timeout 200 python tools/gemm_probe.py "$sh" 8 2>&1 | grep -v Warning | tail -n 1
  done
done
timeout 600 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2k_e2e.log 2>&1; echo "e2e exit $?"; tail -n 4 gpurun_out/r2k_e2e.log
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7
