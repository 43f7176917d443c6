#!/bin/bash
# multi-slot barrier rounds in the MMA issuers: correctness under the guarded build first, then timings
mkdir -p gpurun_out
G=$PWD/comfyui-egregora-audio-super-resolution_b200/libegregora_b200_guard.so
EGREGORA_B200_LIB=$G timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2r_guard.log 2>&1
rc=$?; echo "guarded ops+mega exit $rc"; tail -n 5 gpurun_out/r2r_guard.log
[ $rc -ne 0 ] && exit 1
EGREGORA_B200_LIB=$G EGR_TC_FORCE_PAIR=1 timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2r_guard_pair.log 2>&1
rc=$?; echo "guarded ops (forced pair) exit $rc"; tail -n 3 gpurun_out/r2r_guard_pair.log
[ $rc -ne 0 ] && exit 1
out=gpurun_out/r2r_probe.txt; : > $out
for g in 1 2 4; do
  for b in 1 8; do
    echo "== EGR_TC_GMAX=$g batch $b" >> $out
    EGR_TC_GMAX=$g timeout 300 python tools/gemm_probe.py "" $b 2>&1 | grep -v Warning | grep "conv" >> $out
  done
done
echo "== EGR_TC_GMAX=4 EGR_TC_NO_PAIR=1 batch 8" >> $out
EGR_TC_NO_PAIR=1 timeout 300 python tools/gemm_probe.py "conv2d" 8 2>&1 | grep "conv" >> $out
cat $out
for g in 1 4; do EGR_TC_GMAX=$g timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -6; done
for g in 1 4; do EGR_TC_GMAX=$g timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -6; done
