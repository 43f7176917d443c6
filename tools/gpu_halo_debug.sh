#!/bin/bash
# halo-mode descriptor experiments: same tests under both base-offset encodings and with halo disabled
mkdir -p gpurun_out
for mode in 0 1; do
  echo "=== BASEOFF=$mode"
  EGR_TC_BASEOFF=$mode timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k "stride2 or transposed or conv1d or conv_transpose or attention_gemm" 2>&1 | tail -n 8
done
echo "=== NO_HALO"
EGR_TC_NO_HALO=1 timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k "stride2 or transposed or conv1d or conv_transpose or attention_gemm" 2>&1 | tail -n 4
