#!/bin/bash
# round 2, session 2, sixth call: residual L2 prefetch from the A producer
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2jj_ops.log 2>&1; rc=$?; echo "ops exit $rc"; tail -n 3 gpurun_out/r2jj_ops.log
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/op_times.py 1 > gpurun_out/r2jj_ops_b1.tsv 2>/dev/null
timeout 300 python tools/op_times.py 8 > gpurun_out/r2jj_ops_b8.tsv 2>/dev/null
EGR_TC_NO_PF_RESID=1 timeout 300 python tools/op_times.py 1 > gpurun_out/r2jj_ops_b1_nopf.tsv 2>/dev/null
EGR_TC_NO_PF_RESID=1 timeout 300 python tools/op_times.py 8 > gpurun_out/r2jj_ops_b8_nopf.tsv 2>/dev/null
for f in gpurun_out/r2jj_ops_b1.tsv gpurun_out/r2jj_ops_b1_nopf.tsv gpurun_out/r2jj_ops_b8.tsv gpurun_out/r2jj_ops_b8_nopf.tsv; do
  echo "$f: tc us $(awk -F'\t' '$3=="tc" {s+=$8} END {print s}' $f) convs2 us $(grep 'convs2' $f | awk -F'\t' '{s+=$8} END {print s}') convs1 us $(grep 'convs1' $f | awk -F'\t' '{s+=$8} END {print s}') vae conv2 us $(grep 'vae.*conv2' $f | awk -F'\t' '{s+=$8} END {print s}') total us $(awk -F'\t' '{s+=$8} END {print s}' $f)"
done
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -6
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -6
export EGREGORA_B200_LIB=$PWD/comfyui-egregora-audio-super-resolution_b200/libegregora_b200_trace.so
timeout 300 python tools/gemm_trace.py "conv1d 48->48 k3" 1 brief 2>&1 | tail -3
timeout 300 python tools/gemm_trace.py "conv1d 48->48 k3" 8 brief 2>&1 | tail -3
