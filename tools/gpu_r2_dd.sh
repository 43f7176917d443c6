#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2dd_ops.log 2>&1; rc=$?; echo "ops+mega exit $rc"; tail -n 3 gpurun_out/r2dd_ops.log
[ $rc -ne 0 ] && exit 1
timeout 300 python tools/op_times.py 1 > gpurun_out/r2dd_ops_b1.tsv 2>/dev/null
grep -i "norm" gpurun_out/r2dd_ops_b1.tsv | awk -F'\t' '{s+=$8} END {print "GN total us", s}'
grep -i "norm" profiles/r2_g_ops_b1.tsv | awk -F'\t' '{s+=$8} END {print "GN total us before", s}'
grep "vae.encoder.down.0.block.0.norm1\|vae.decoder.up.0.block.0.norm1" gpurun_out/r2dd_ops_b1.tsv
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -6
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -6
timeout 900 python -m pytest tests/test_flashsr_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2dd_e2e.log 2>&1; echo "e2e exit $?"; tail -n 3 gpurun_out/r2dd_e2e.log
