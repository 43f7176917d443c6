#!/bin/bash
# round 2, session 2, fifth call: chunked N=1 conv kernel + clock trace of the 48-channel vocoder convs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "one_output or simt or strided" > gpurun_out/r2ii_ops.log 2>&1; rc=$?; echo "ops exit $rc"; tail -n 3 gpurun_out/r2ii_ops.log
timeout 300 python tools/op_times.py 1 2>/dev/null | awk -F'\t' '$3=="simt"'
timeout 300 python tools/op_times.py 8 2>/dev/null | awk -F'\t' '$3=="simt"'
export EGREGORA_B200_LIB=$PWD/comfyui-egregora-audio-super-resolution_b200/libegregora_b200_trace.so
for b in 1 8; do
  timeout 300 python tools/gemm_trace.py "conv1d 48->48 k3" $b 2>&1 | tail -12
  timeout 300 python tools/gemm_trace.py "conv1d 48->48 k11" $b 2>&1 | tail -12
done
timeout 300 python tools/gemm_trace.py "conv1d 96->96 k3" 1 2>&1 | tail -12
