#!/bin/bash
# bring-up of the CTA-pair GEMM (cta_group::2): library built with GUARD=1 (endless barrier waits trap), every step under a timeout
mkdir -p gpurun_out
EGR_TC_FORCE_PAIR=1 timeout 400 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2h_ops_pair.log 2>&1; echo "ops (forced pair) exit $?"; tail -n 25 gpurun_out/r2h_ops_pair.log
timeout 600 python -m pytest tests/test_flashsr_gpu.py tests/test_mega_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2h_e2e.log 2>&1; echo "e2e exit $?"; tail -n 12 gpurun_out/r2h_e2e.log
cat gpurun_out/parity_full.json
timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7
EGR_TC_NO_PAIR=1 timeout 300 python tools/section_times.py 1 1 2>/dev/null | tail -7
EGR_TC_NO_PAIR=1 timeout 300 python tools/section_times.py 8 1 2>/dev/null | tail -7
