#!/bin/bash
# round 2, session 2, fourth call: GN stats mapping, N=1 conv kernel, vector stores of the 1-channel-input convs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2hh_ops.log 2>&1; rc=$?; echo "ops+mega exit $rc"; tail -n 3 gpurun_out/r2hh_ops.log
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/op_times.py 1 > gpurun_out/r2hh_ops_b1.tsv 2>/dev/null
timeout 300 python tools/op_times.py 8 > gpurun_out/r2hh_ops_b8.tsv 2>/dev/null
for f in gpurun_out/r2hh_ops_b1.tsv profiles/r2_h_ops_b1.tsv gpurun_out/r2hh_ops_b8.tsv profiles/r2_h_ops_b8.tsv; do
  echo "$f: GN us $(grep -i 'norm' $f | awk -F'\t' '{s+=$8} END {print s}') (stats $(grep -i 'norm.*stats' $f | awk -F'\t' '{s+=$8} END {print s}'))  snake us $(grep -i 'activation' $f | awk -F'\t' '{s+=$8} END {print s}') simt us $(awk -F'\t' '$3=="simt" {s+=$8} END {print s}' $f) total us $(awk -F'\t' '{s+=$8} END {print s}' $f)"
done
awk -F'\t' '$3=="simt"' gpurun_out/r2hh_ops_b1.tsv gpurun_out/r2hh_ops_b8.tsv
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -6
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -6
timeout 900 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2hh_e2e.log 2>&1; echo "e2e exit $?"; tail -n 3 gpurun_out/r2hh_e2e.log
cat gpurun_out/parity_full.json
