#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2aa_ops.log 2>&1; rc=$?; echo "ops+mega exit $rc"; tail -n 3 gpurun_out/r2aa_ops.log
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/op_times.py 1 > gpurun_out/r2aa_ops_b1.tsv 2>/dev/null
timeout 300 python tools/op_times.py 8 > gpurun_out/r2aa_ops_b8.tsv 2>/dev/null
python tools/op_diff.py -v gpurun_out/r2y_ops_hg1.tsv gpurun_out/r2aa_ops_b1.tsv
python tools/op_diff.py -v gpurun_out/r2z_ops_b8_skip0.tsv gpurun_out/r2aa_ops_b8.tsv
grep "vocoder.ups" gpurun_out/r2aa_ops_b1.tsv gpurun_out/r2aa_ops_b8.tsv
timeout 900 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2aa_e2e.log 2>&1; echo "e2e exit $?"; tail -n 4 gpurun_out/r2aa_e2e.log
